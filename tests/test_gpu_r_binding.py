"""GPU: the R-side binding bindings/R/atlasqtl_b200_core.R -- the reference package's VB outer loop written in R against
the stateful C ABI -- executed end to end on the CUDA library: the R evaluator of oracle/rlite runs the R file, its
`.Call(`_atlasqtl_aq_*`)` go through an emulation of bindings/R/atlasqtl_b200_shim.c (same symbols, arguments, checks,
returned lists) to `libatlasqtl_b200.so` via ctypes.  Expected values: outputs of the reference's own R code
(tests/golden/rlite_core_*.npz).  Bars: identical iteration count, ELBO <= 1e-10 relative at every evaluation,
gam_vb / beta_vb <= 1e-8, identical {gam_vb > 0.5}."""
import numpy as np
import pytest

from rlite_cases import CORE_FILES, CORE_IDS, load_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", CORE_FILES, ids=CORE_IDS)
def test_r_binding_on_the_cuda_library(path):
    import r_binding
    from atlasqtl_b200.device import SweepContext
    g, hyper, init, anneal = load_case(path)
    it, shim = r_binding.load(lambda X, Y: SweepContext(X, Y), helpers="standin")
    trace = []
    out = r_binding.run_core(it, g["Y"], g["X"], anneal, float(g["tol"]), hyper, init, thinned=bool(g["thinned"]),
                             trace=trace)
    lb = np.array([v for _, v in trace])
    assert bool(out["converged"][0]) and int(out["it"][0]) == int(g["it"])
    assert lb.shape == g["lb"].shape
    assert np.max(np.abs(lb - g["lb"]) / np.abs(g["lb"])) <= 1e-10
    assert np.abs(out["gam_vb"] - g["gam_vb"]).max() <= 1e-8
    assert np.abs(out["beta_vb"] - g["beta_vb"]).max() <= 1e-8
    np.testing.assert_allclose(out["theta_vb"], g["theta_vb"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(out["zeta_vb"], g["zeta_vb"], rtol=1e-8, atol=1e-9)
    assert np.array_equal(out["gam_vb"] > 0.5, g["gam_vb"] > 0.5)
    assert shim.live == 0
