"""CPU: the R-side binding bindings/R/atlasqtl_b200_core.R (the VB outer loop of the R package written against the
stateful C ABI) is EXECUTED -- by the R evaluator of oracle/rlite, `.Call` going to an emulation of the C shim over the
oracle-backed test double of the device -- and must reproduce the outputs of the reference's own R code
(tests/golden/rlite_core_*.npz): same iteration count, ELBO at every evaluation, parameters.

helpers = "reference": the package helpers the binding calls unchanged come from the reference's own R files (only where
/root/reference exists); helpers = "standin": NumPy stand-ins, the configuration the GPU box runs
(tests/test_gpu_r_binding.py)."""
import numpy as np
import pytest

import r_binding
from fake_context import OracleSweepContext
from rlite_cases import CORE_FILES, CORE_IDS, load_case


def _check(out, trace, g, tol_lb=1e-10, tol_par=1e-9):
    lb = np.array([v for _, v in trace])
    assert bool(out["converged"][0]) and int(out["it"][0]) == int(g["it"])
    assert lb.shape == g["lb"].shape
    assert np.max(np.abs(lb - g["lb"]) / np.abs(g["lb"])) <= tol_lb
    assert abs(float(out["lb_opt"][0]) - float(g["lb_opt"])) <= tol_lb * abs(float(g["lb_opt"]))
    assert abs(float(out["diff_lb"][0]) - float(g["diff_lb"])) <= 1e-6
    for k in ("gam_vb", "beta_vb", "theta_vb", "zeta_vb"):
        assert np.max(np.abs(out[k] - g[k])) <= tol_par, k
    assert np.array_equal(out["gam_vb"] > 0.5, g["gam_vb"] > 0.5)


@pytest.mark.parametrize("helpers", ["standin", "reference"])
@pytest.mark.parametrize("path", CORE_FILES, ids=CORE_IDS)
def test_r_binding_reproduces_the_reference_r_code(oracle_built, path, helpers):
    if helpers == "reference":
        from oracle.rlite import reference as R
        if not R.available():
            pytest.skip("/root/reference is not present here")
    g, hyper, init, anneal = load_case(path)
    if g["X"].shape[1] * g["Y"].shape[1] > 20000:
        pytest.skip("covered by the smaller cases; the test double is slow")
    it, shim = r_binding.load(lambda X, Y: OracleSweepContext(X, Y), helpers=helpers)
    trace = []
    out = r_binding.run_core(it, g["Y"], g["X"], anneal, float(g["tol"]), hyper, init, thinned=bool(g["thinned"]),
                             trace=trace)
    _check(out, trace, g)
    assert shim.live == 0 and shim.calls["_atlasqtl_aq_create"] == 1          # the context is destroyed at the end
    sweeps = shim.calls.get("_atlasqtl_aq_sweep", 0) + shim.calls.get("_atlasqtl_aq_sweep_mis", 0)
    assert sweeps == int(g["it"]) and shim.calls["_atlasqtl_aq_rowsums_zpart"] == int(g["it"])
    assert shim.calls["_atlasqtl_aq_get_state"] == 1                          # p x q objects cross the boundary once
    assert not it.warnings


def test_r_binding_full_output_and_argument_checks(oracle_built):
    g, hyper, init, anneal = load_case(CORE_FILES[1])
    it, shim = r_binding.load(lambda X, Y: OracleSweepContext(X, Y))
    out = r_binding.run_core(it, g["Y"], g["X"], anneal, float(g["tol"]), hyper, init, full_output=True)
    for k in ("eta_vb", "kappa_vb", "lam2_inv_vb", "nu_s0_vb", "nu_vb", "rho_s0_vb", "rho_vb", "rho_xi_inv_vb",
              "sig02_inv_vb", "sig2_inv_vb", "sig2_theta_vb", "sig2_zeta_vb", "tau_vb", "xi_inv_vb"):
        np.testing.assert_allclose(np.asarray(out[k]).reshape(g["full_" + k].shape), g["full_" + k], rtol=1e-8, err_msg=k)
    # the emulated shim checks its arguments like the C one: an R-side slip fails loudly instead of reading out of bounds
    with pytest.raises(Exception, match="double vector of length"):
        it.run("ctx <- .Call(`_atlasqtl_aq_create`, matrix(0.5, 4, 3), matrix(0.5, 4, 2), 0L); "
               ".Call(`_atlasqtl_aq_refresh_tables`, ctx, c(0, 0), c(0, 0), 1, FALSE)")
    with pytest.raises(Exception, match="double matrix"):
        it.run("ctx <- .Call(`_atlasqtl_aq_create`, matrix(0.5, 4, 3), matrix(0.5, 4, 2), 0L); "
               ".Call(`_atlasqtl_aq_set_state`, ctx, matrix(0, 3, 3), matrix(0, 3, 2))")
    with pytest.raises(Exception, match="Batch scheme"):
        r_binding.load(lambda X, Y: OracleSweepContext(X, Y))[0].run(
            "atlasqtl_b200_core_(matrix(0, 2, 2), matrix(0, 2, 2), 2, NULL, 1, 0.1, 10, 0, list(), list(), batch = '0')")


def test_r_binding_checkpoints_through_the_package_checkpoint_function(oracle_built, tmp_path):
    """checkpoint_ / checkpoint_clean_up_ (R/utils.R:571-627) are the reference's own functions here: every 100th
    non-annealed iteration the binding fetches gam_vb / beta_vb from the device (aq_get_state, two separately allocated
    matrices) and hands them over; the files are removed at the end.  The case runs exactly 100 iterations."""
    from oracle.rlite import reference as R
    from oracle.rlite.values import Builtin, from_py, lgl
    if not R.available():
        pytest.skip("/root/reference is not present here")
    g, hyper, init, anneal = load_case([f for f in CORE_FILES if "d_linear" in f][0])
    assert int(g["it"]) == 100
    it, shim = r_binding.load(lambda X, Y: OracleSweepContext(X, Y), helpers="reference")
    seen = []
    orig = it.globalenv.vars["checkpoint_clean_up_"]

    def spy(it_, pos, named):
        import glob
        for f in sorted(glob.glob(str(tmp_path) + "/tmp_output_it_*.RData")):
            seen.append((f, {k: v.copy() for k, v in np.load(f).items()}))
        return it_.apply(orig, pos, named, it_.globalenv)
    it.globalenv.vars["checkpoint_clean_up_"] = Builtin(spy, "checkpoint_clean_up_")
    q = g["Y"].shape[1]
    from oracle.rlite.values import RList
    out = it.call("atlasqtl_b200_core_", from_py(np.array(g["Y"], order="F")), from_py(np.array(g["X"], order="F")),
                  from_py(float(q)), from_py(np.asarray(anneal, dtype=np.float64)), from_py(1.0), from_py(float(g["tol"])),
                  from_py(1000.0), from_py(0.0),
                  RList([from_py(np.asarray(v, dtype=np.float64)) for v in hyper.values()], list(hyper.keys())),
                  RList([from_py(np.array(v, dtype=np.float64, order="F")) for v in init.values()], list(init.keys())),
                  checkpoint_path=from_py(str(tmp_path) + "/"), debug=lgl(True))
    assert len(seen) == 1 and seen[0][0].endswith("tmp_output_it_100.RData")
    ck = seen[0][1]
    assert int(ck["tmp_vb$it"][0]) == 100 and bool(ck["tmp_vb$converged"][0])
    np.testing.assert_allclose(ck["tmp_vb$gam_vb"], g["gam_vb"], atol=1e-9)       # the state of iteration 100 = the final one
    np.testing.assert_allclose(ck["tmp_vb$beta_vb"], g["beta_vb"], atol=1e-9)
    assert not np.shares_memory(ck["tmp_vb$gam_vb"], ck["tmp_vb$beta_vb"]) and np.abs(ck["tmp_vb$gam_vb"] - ck["tmp_vb$beta_vb"]).max() > 0
    import glob
    assert glob.glob(str(tmp_path) + "/tmp_output_it_*") == []                      # checkpoint_clean_up_ removed it
    assert shim.calls["_atlasqtl_aq_get_state"] == 2
    assert np.abs(np.asarray(out.get("gam_vb").a) - g["gam_vb"]).max() <= 1e-9
