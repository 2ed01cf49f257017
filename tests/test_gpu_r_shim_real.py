"""GPU: the R host loop (bindings/R/atlasqtl_b200_core.R, run by the R evaluator of oracle/rlite) through the SHIPPED C
shim (bindings/R/atlasqtl_b200_shim.c, compiled against the stub R runtime of tests/r_stub) into libatlasqtl_b200.so: the
complete R-facing stack of INTEGRATION.md section 2b, nothing emulated but R itself.  Expected values: outputs of the
reference's own R code (tests/golden/rlite_core_*.npz); the stateless 15-argument entry against the reference's compiled
loop."""
import numpy as np
import pytest

from rlite_cases import CORE_FILES, CORE_IDS, load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def shim_library():
    """The shim + stub runtime are compiled with gcc at test time; a box without a C compiler skips (the build itself is
    covered by the CPU suite, tests/test_r_shim_real.py)."""
    import r_shim_real
    try:
        r_shim_real.load_library()
    except (RuntimeError, OSError) as e:
        pytest.skip(f"cannot build / load the shim against the stub R runtime here: {e}")


@pytest.mark.parametrize("path", CORE_FILES, ids=CORE_IDS)
def test_r_loop_through_the_real_shim_on_the_cuda_library(path):
    import r_binding
    g, hyper, init, anneal = load_case(path)
    it, shim = r_binding.load(helpers="standin", shim="real")
    trace = []
    out = r_binding.run_core(it, g["Y"], g["X"], anneal, float(g["tol"]), hyper, init, thinned=bool(g["thinned"]),
                             trace=trace)
    lb = np.array([v for _, v in trace])
    assert bool(out["converged"][0]) and int(out["it"][0]) == int(g["it"])
    assert lb.shape == g["lb"].shape
    assert np.max(np.abs(lb - g["lb"]) / np.abs(g["lb"])) <= 1e-10
    assert np.abs(out["gam_vb"] - g["gam_vb"]).max() <= 1e-8
    assert np.abs(out["beta_vb"] - g["beta_vb"]).max() <= 1e-8
    assert np.array_equal(out["gam_vb"] > 0.5, g["gam_vb"] > 0.5)
    sweeps = shim.calls.get("_atlasqtl_aq_sweep", 0) + shim.calls.get("_atlasqtl_aq_sweep_mis", 0)
    assert sweeps == int(g["it"]) and shim.calls["_atlasqtl_aq_destroy"] == 1
    assert not it.warnings


def test_stateless_entry_through_the_real_shim(oracle_built):
    """coreDualLoop(...) as the R closure of the reference (R/RcppExports.R:4-6): 15 SEXPs, in-place outputs."""
    from oracle import native
    from oracle.rlite.interp import Interp
    from oracle.rlite.values import Builtin, chr_, from_py
    from problems import make_problem, sweep_inputs
    from r_shim_real import RealShim
    if not native.ref_available():
        pytest.skip("oracle/_ref is not available on this box")
    X, Y, hyper, init = make_problem(120, 40, 9)
    p, q = X.shape[1], Y.shape[1]
    si = sweep_inputs(X, Y, init, c=0.8)
    cp_X, cp_Y_X = np.asfortranarray(X.T @ X), np.asfortranarray(Y.T @ X)
    beta0 = np.asfortranarray(si["gam"] * si["mu"])
    cbx0 = np.asfortranarray(cp_X @ beta0)
    order = np.random.default_rng(1).permutation(p).astype(np.int32)
    sq = np.arange(q, dtype=np.int32)
    ref = [a.copy(order="F") for a in (si["gam"], beta0, cbx0, si["mu"])]
    native.core_dual_loop(cp_X, cp_Y_X, ref[0], si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"], si["log_tau"], ref[1],
                          ref[2], ref[3], si["sig2_beta"], si["tau"], order, sq, c=0.8, impl="reference")
    it = Interp()
    it.globalenv.vars[".Call"] = Builtin(RealShim(), ".Call")
    it.globalenv.vars["_atlasqtl_coreDualLoop"] = chr_("_atlasqtl_coreDualLoop")
    it.run("coreDualLoop <- function(cp_X, cp_Y_X, gam_vb, lp, l1p, log_sig2_inv_vb, log_tau_vb, m1_beta, cp_betaX_X, "
           "mu_beta_vb, sig2_beta_vb, tau_vb, shuffled_ind, sample_q, c = 1) invisible(.Call(`_atlasqtl_coreDualLoop`, "
           "cp_X, cp_Y_X, gam_vb, lp, l1p, log_sig2_inv_vb, log_tau_vb, m1_beta, cp_betaX_X, mu_beta_vb, sig2_beta_vb, "
           "tau_vb, shuffled_ind, sample_q, c))")
    mine = [from_py(a.copy(order="F")) for a in (si["gam"], beta0, cbx0, si["mu"])]
    it.call("coreDualLoop", from_py(cp_X), from_py(cp_Y_X), mine[0], from_py(si["log_Phi"]), from_py(si["log_1_min_Phi"]),
            from_py(float(si["log_sig2_inv"])), from_py(si["log_tau"]), mine[1], mine[2], mine[3], from_py(si["sig2_beta"]),
            from_py(si["tau"]), from_py(order.astype(np.int64)), from_py(sq.astype(np.int64)), c=from_py(0.8))
    for a, b, tol in zip(mine, ref, (1e-8, 1e-8, 1e-7, 1e-8)):
        assert np.abs(a.a - b).max() <= tol          # in place on the evaluator's own matrices, as under R
