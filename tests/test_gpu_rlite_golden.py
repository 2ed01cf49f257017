"""GPU: the CUDA path against outputs of the reference's OWN R code.

tests/golden/rlite_*.npz were produced by running the reference's unmodified R sources (outer loop
R/atlasqtl_global_local_core.R:8-433 with update_vb.R / elbo.R / utils.R, `atlasqtl()` R/atlasqtl.R:179-322 with
prepare_data_) through the R evaluator of oracle/rlite, `.Call` bound to the reference's own src/coreLoop.cpp
(tests/golden/make_rlite_golden.py; inputs are stored in the files).  The box has no /root/reference and needs none.

Bars: identical iteration count, ELBO relative <= 1e-10 at every evaluation, max|d gam_vb| and
max|d beta_vb| <= 1e-8, theta / zeta to 1e-8, identical {gam_vb > 0.5} and {bFDR < 0.05} sets.
"""
import os

import numpy as np
import pytest

from rlite_cases import CORE_FILES, CORE_IDS, GOLD, hyper_init_of, load_case

pytestmark = pytest.mark.gpu


def _compare(out, trace, g):
    from oracle import vb_oracle
    lb = np.array([r["lb"] for r in trace if r["lb"] is not None])
    assert out["converged"] and bool(g["converged"])
    assert out["it"] == int(g["it"])
    assert lb.shape == g["lb"].shape
    rel = np.abs(lb - g["lb"]) / np.abs(g["lb"])
    assert rel.max() <= 1e-10, rel.max()
    assert abs(out["lb_opt"] - float(g["lb_opt"])) <= 1e-10 * abs(float(g["lb_opt"]))
    assert np.abs(out["gam_vb"] - g["gam_vb"]).max() <= 1e-8
    assert np.abs(out["beta_vb"] - g["beta_vb"]).max() <= 1e-8
    np.testing.assert_allclose(out["theta_vb"], g["theta_vb"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(out["zeta_vb"], g["zeta_vb"], rtol=1e-8, atol=1e-9)
    assert np.array_equal(out["gam_vb"] > 0.5, g["gam_vb"] > 0.5)
    assert np.array_equal(vb_oracle.assign_bFDR(out["gam_vb"]) < 0.05, vb_oracle.assign_bFDR(g["gam_vb"]) < 0.05)


@pytest.mark.parametrize("path", CORE_FILES, ids=CORE_IDS)
def test_cuda_run_reproduces_the_reference_r_code(path):
    from atlasqtl_b200 import core
    g, hyper, init, anneal = load_case(path)
    X, Y = g["X"], g["Y"]
    trace = []
    out = core.atlasqtl_global_local_core_(Y, X, Y.shape[1], anneal, 1, float(g["tol"]), 1000, 0, hyper, init,
                                           thinned_elbo_eval=bool(g["thinned"]), debug=True, trace=trace)
    _compare(out, trace, g)


@pytest.mark.parametrize("prepare_on_device", [False, True])
def test_cuda_atlasqtl_call_reproduces_the_reference_r_code(prepare_on_device):
    """atlasqtl() on raw calls with constant / duplicated columns and missing responses, the pre-processing on the host
    mirror or on the device."""
    from atlasqtl_b200 import atlasqtl
    g = np.load(os.path.join(GOLD, "rlite_atlasqtl_top.npz"))
    hyper, init = hyper_init_of(g)
    trace = []
    out = atlasqtl(g["Y_raw"], g["X_raw"], None, anneal=tuple(g["anneal"]), tol=float(g["tol"]), maxit=1000, verbose=0,
                   list_hyper=hyper, list_init=init, trace=trace, prepare_on_device=prepare_on_device)
    assert out["rmvd_cst_x"] == list(g["prep_rmvd_cst_x"])
    pairs = sorted((k, r) for k, rs in out["rmvd_coll_x"].items() for r in rs)
    assert pairs == sorted(zip(g["prep_rmvd_coll_kept"].tolist(), g["prep_rmvd_coll_x"].tolist()))
    assert out["names_x"] == list(g["names_x"]) and out["names_y"] == list(g["names_y"])
    _compare(out, trace, g)
