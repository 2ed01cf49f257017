"""CPU: pin the oracle.  (i) cavi_oracle.c against the golden vectors produced by the reference's own
coreDualLoop (tests/golden, bit-exact) and, where oracle/_ref is present, against that library live;
(ii) dual == primal == blocked forms; (iii) the restated R outer loop obeys the reference's own invariants."""
import glob
import os

import numpy as np
import pytest

from problems import make_problem, sweep_inputs

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "coredualloop_*.npz")))


def _dual_inputs(d):
    X, Y = d["X"], d["Y"]
    gam, mu = np.array(d["gam"], order="F"), np.array(d["mu"], order="F")
    beta = np.asfortranarray(gam * mu)
    cp_X, cp_Y_X = np.asfortranarray(X.T @ X), np.asfortranarray(Y.T @ X)
    return gam, mu, beta, cp_X, cp_Y_X, np.asfortranarray(cp_X @ beta)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[14:-4] for p in GOLDEN])
def test_dual_oracle_matches_reference_golden_bit_exact(oracle_built, path):
    native = oracle_built
    d = dict(np.load(path))
    q = d["Y"].shape[1]
    gam, mu, beta, cp_X, cp_Y_X, cbx = _dual_inputs(d)
    native.core_dual_loop(cp_X, cp_Y_X, gam, np.asfortranarray(d["log_Phi"]), np.asfortranarray(d["log_1_min_Phi"]),
                          float(d["log_sig2_inv"]), d["log_tau"], beta, cbx, mu, d["sig2_beta"], d["tau"], d["order"],
                          np.arange(q, dtype=np.int32), c=float(d["c"]), impl="oracle")
    assert np.array_equal(gam, d["out_gam"])
    assert np.array_equal(mu, d["out_mu"])
    assert np.array_equal(beta, d["out_beta"])
    assert np.array_equal(cbx, d["out_cp_betaX_X"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[14:-4] for p in GOLDEN])
@pytest.mark.parametrize("form", ["primal", "blocked"])
def test_sample_space_forms_match_golden(oracle_built, path, form):
    native = oracle_built
    d = dict(np.load(path))
    X, Y = np.asfortranarray(d["X"]), np.asfortranarray(d["Y"])
    gam, mu = np.array(d["gam"], order="F"), np.array(d["mu"], order="F")
    beta = np.asfortranarray(gam * mu)
    R = native.residual(X, Y, beta)
    args = (gam, np.asfortranarray(d["log_Phi"]), np.asfortranarray(d["log_1_min_Phi"]), float(d["log_sig2_inv"]),
            d["log_tau"], beta, mu, d["sig2_beta"], d["tau"], d["order"])
    if form == "primal":
        native.sweep_primal(X, np.asfortranarray((X ** 2).sum(0)), R, *args, c=float(d["c"]), nthreads=3)
    else:
        native.sweep_primal_blocked(X, R, *args, c=float(d["c"]), B=8)
    assert np.abs(gam - d["out_gam"]).max() <= 1e-12
    assert np.abs(mu - d["out_mu"]).max() <= 1e-12
    # the residual carries the reference's running X'X beta:  X'(Y - R) == cp_betaX_X
    np.testing.assert_allclose(X.T @ (Y - R), d["out_cp_betaX_X"], atol=1e-9)


def test_dual_oracle_matches_live_reference_library(oracle_built):
    native = oracle_built
    if not native.ref_available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    X, Y, hyper, init = make_problem(120, 90, 25)
    si = sweep_inputs(X, Y, init, c=0.8)
    order = np.random.default_rng(1).permutation(X.shape[1]).astype(np.int32)
    outs = {}
    for impl in ("oracle", "reference"):
        d = dict(X=X, Y=Y, gam=si["gam"], mu=si["mu"])
        gam, mu, beta, cp_X, cp_Y_X, cbx = _dual_inputs(d)
        native.core_dual_loop(cp_X, cp_Y_X, gam, si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"], si["log_tau"],
                              beta, cbx, mu, si["sig2_beta"], si["tau"], order, np.arange(Y.shape[1], dtype=np.int32),
                              c=0.8, impl=impl)
        outs[impl] = (gam, mu, beta, cbx)
    for a, b in zip(outs["oracle"], outs["reference"]):
        assert np.array_equal(a, b)


def test_sample_q_subset_only_touches_its_columns(oracle_built):
    native = oracle_built
    X, Y, hyper, init = make_problem(60, 30, 10)
    si = sweep_inputs(X, Y, init)
    d = dict(X=X, Y=Y, gam=si["gam"], mu=si["mu"])
    gam, mu, beta, cp_X, cp_Y_X, cbx = _dual_inputs(d)
    g0 = gam.copy()
    native.core_dual_loop(cp_X, cp_Y_X, gam, si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"], si["log_tau"], beta,
                          cbx, mu, si["sig2_beta"], si["tau"], np.arange(X.shape[1], dtype=np.int32),
                          np.array([2, 7], dtype=np.int32), c=1.0)
    changed = np.flatnonzero(np.any(gam != g0, axis=0))
    assert list(changed) == [2, 7]


@pytest.mark.parametrize("anneal", [None, (1, 2, 10)])
def test_restated_r_loop_invariants(oracle_built, anneal):
    """tests/testthat/main.R recipe (n=100, p=75, q=20, p0=c(5,25)): converges (test_convergence.R:5-7) and the
    ELBO never decreases after annealing (debug stop, R/atlasqtl_global_local_core.R:359-360)."""
    from oracle import vb_oracle
    X, Y, hyper, init = make_problem(100, 75, 20, p_act=10, q_act=20, maf=0.2, p0=(5, 25))
    tr = []
    out = vb_oracle.atlasqtl_global_local_core_(Y, X, Y.shape[1], anneal, 1, 0.1, 1000, hyper, init, sweep="dual",
                                                debug=True, trace=tr)
    assert out["converged"]
    lbs = [r["lb"] for r in tr if r["lb"] is not None]
    assert all(b + 1.49e-8 >= a for a, b in zip(lbs, lbs[1:]))
    if anneal is not None:
        assert all(r["lb"] is None for r in tr[: anneal[2] - 1])  # no ELBO while annealing
        np.testing.assert_allclose([r["c"] for r in tr[:10]], 2.0 ** (-(9 - np.arange(10)) / 9), rtol=1e-14)


def test_q_approx_vec_and_bfdr():
    from scipy import integrate

    from oracle import vb_oracle
    x = np.array([0.05, 0.7, 1.0, 1.3, 4.0, 60.0])
    ref = [integrate.quad(lambda t: np.exp(-xx * t) / (1 + t), 0, np.inf)[0] for xx in x]  # E1(x) e^x
    np.testing.assert_allclose(vb_oracle.Q_approx_vec(x), ref, rtol=2e-7)
    ppi = np.array([[0.9, 0.2], [0.99, 0.6]])
    fdr = vb_oracle.assign_bFDR(ppi)
    np.testing.assert_allclose(fdr, [[(0.01 + 0.1) / 2, (0.01 + 0.1 + 0.4 + 0.8) / 4], [0.01, (0.01 + 0.1 + 0.4) / 3]])


def test_masked_primal_sweep_matches_reference_mis_loop(oracle_built):
    """coreDualMisLoop (src/coreLoop.cpp:91-138) with the per-trait Gram corrections cp_X_rm of
    R/atlasqtl_global_local_core.R:25-32, against the sample-space restatement with a masked residual."""
    from problems import mis_inputs
    native = oracle_built
    if not native.ref_available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    X, Y, hyper, init = make_problem(60, 40, 12)
    p, q = X.shape[1], Y.shape[1]
    si = sweep_inputs(X, Y, init, c=0.7)
    mi = mis_inputs(X, Y, si)
    order = np.random.default_rng(2).permutation(p).astype(np.int32)
    # the reference's inputs
    cp_X = np.asfortranarray(X.T @ X)
    cp_X_rm = np.zeros((p, p, q), order="F")
    for k in range(q):
        rows = np.flatnonzero(mi["mis"][:, k] == 0)
        cp_X_rm[:, :, k] = X[rows].T @ X[rows]
    cp_Y_X = np.asfortranarray(mi["Y"].T @ X)
    gam_r, mu_r = si["gam"].copy(order="F"), si["mu"].copy(order="F")
    beta_r = np.asfortranarray(gam_r * mu_r)
    cbx = np.asfortranarray(cp_X @ beta_r - np.stack([cp_X_rm[:, :, k] @ beta_r[:, k] for k in range(q)], axis=1))
    native.ref_core_dual_mis_loop(cp_X, cp_X_rm, cp_Y_X, gam_r, si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"],
                                  si["log_tau"], beta_r, cbx, mu_r, mi["sig2_beta"], si["tau"], order,
                                  np.arange(q, dtype=np.int32), c=0.7)
    # ours
    gam, mu = si["gam"].copy(order="F"), si["mu"].copy(order="F")
    beta = np.asfortranarray(gam * mu)
    R = np.asfortranarray(mi["mis"] * (mi["Y"] - X @ beta))
    native.sweep_primal_mis(X, mi["mis"], mi["xnsq"], R, gam, si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"],
                            si["log_tau"], beta, mu, mi["sig2_beta"], si["tau"], order, c=0.7)
    assert np.abs(gam - gam_r).max() <= 1e-12
    assert np.abs(mu - mu_r).max() <= 1e-12
    assert np.abs(beta - beta_r).max() <= 1e-12
    # the reference's running X'(mis o X) beta equals X'(mis o Y - R)
    np.testing.assert_allclose(X.T @ (mi["Y"] - R), cbx, atol=1e-9)


def test_prepare_oracle_and_host_mirror_agree():
    """The literal restatement of R's scale / rm_constant_ / rm_collinear_ (oracle/prepare_oracle.py) and the package's
    vectorised host mirror (atlasqtl_b200/prepare.py) are two independent computations of prepare_data_."""
    from atlasqtl_b200 import prepare
    from oracle import prepare_oracle
    rng = np.random.default_rng(8)
    G = rng.binomial(2, 0.3, size=(80, 40)).astype(np.float64)
    G[:, 1::9] = rng.normal(1.0, 3.0, size=(80, len(range(1, 40, 9))))
    G[:, 4] = 1.0
    G[:, 13] = 0.5
    G[:, 30] = G[:, 7]
    G[:, 31] = G[:, 7]
    G[:, 2] = G[:, 38]
    Y = rng.normal(size=(80, 6)) + 2.0
    Y[rng.uniform(size=Y.shape) < 0.1] = np.nan
    o = prepare_oracle.prepare_data_(Y, G)
    h = prepare.prepare_data_(Y, G, 0.1, 10)
    assert np.array_equal(o["bool_rmvd_x"], h["bool_rmvd_x"])
    assert o["bool_cst_x"].sum() == 2 and o["bool_coll_x"].sum() == 3
    assert list(o["dup_of"][[30, 31, 38]]) == [7, 7, 2]
    assert h["rmvd_coll_x"] == {"Cov_x_8": ["Cov_x_31", "Cov_x_32"], "Cov_x_3": ["Cov_x_39"]}
    np.testing.assert_allclose(o["X"], h["X"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(o["Y"], h["Y"], rtol=0, atol=1e-13, equal_nan=True)
    n = G.shape[0]
    np.testing.assert_allclose((o["X"] ** 2).sum(axis=0), n - 1, rtol=1e-12)   # the sweep's pre-condition


GOLDEN_MIS = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "coredualmisloop_*.npz")))


@pytest.mark.parametrize("path", GOLDEN_MIS, ids=[os.path.basename(p) for p in GOLDEN_MIS])
def test_masked_primal_sweep_matches_reference_mis_golden(oracle_built, path):
    """Committed outputs of the reference's own coreDualMisLoop (tests/golden/make_golden.py) vs the masked-residual
    restatement: the pin holds where oracle/_ref cannot be rebuilt."""
    native = oracle_built
    d = np.load(path)
    X = np.asfortranarray(d["X"])
    mis, Ym = np.asfortranarray(d["mis"]), np.asfortranarray(d["Y_mis"])
    gam, mu = d["gam"].copy(order="F"), d["mu"].copy(order="F")
    beta = np.asfortranarray(gam * mu)
    R = np.asfortranarray(mis * (Ym - X @ beta))
    native.sweep_primal_mis(X, mis, np.asfortranarray(d["xnsq"]), R, gam, np.asfortranarray(d["log_Phi"]),
                            np.asfortranarray(d["log_1_min_Phi"]), float(d["log_sig2_inv"]), d["log_tau"], beta, mu,
                            np.asfortranarray(d["sig2_beta_pq"]), d["tau"], d["order"], c=float(d["c"]))
    assert np.abs(gam - d["out_gam"]).max() <= 1e-12
    assert np.abs(mu - d["out_mu"]).max() <= 1e-12
    assert np.abs(beta - d["out_beta"]).max() <= 1e-12
    np.testing.assert_allclose(X.T @ (Ym - R), d["out_cp_betaX_X"], atol=1e-9)
