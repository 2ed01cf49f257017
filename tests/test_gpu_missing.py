"""GPU: the missing-response path (coreDualMisLoop, reference src/coreLoop.cpp:91-138) through the C ABI against
the reference's own loop (dual form with the per-trait cp_X_rm corrections) and the sample-space oracle."""
import numpy as np
import pytest

from problems import make_problem, mis_inputs, sweep_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tile", "warp"], autouse=True)
def mis_kernel(request, monkeypatch):
    """Both CUDA kernels behind aq_sweep_mis: the tile-structured one (blocked tensor-core sweep with masked accumulators and
    a per-trait Gram band table; the default whenever the table fits) and the warp-per-trait one (no table, n <= 2048)."""
    monkeypatch.setenv("AQ_MIS_KERNEL", request.param)
    return request.param


@pytest.mark.parametrize("n,p,q,c,shuffle", [
    (60, 40, 12, 0.7, True),       # dual-feasible: checked against the reference's coreDualMisLoop itself
    (100, 75, 20, 1.0, False),
    (130, 33, 9, 0.9, True),       # n, p, q off every 32 / 8 boundary
    (700, 50, 17, 1.0, True),      # 32 samples per lane
    (1500, 40, 11, 0.8, True),     # 64 samples per lane
    (784, 27, 35, 0.9, True),      # tile kernel: the largest single-CTA configuration, three tiles
    (2500, 30, 20, 1.0, True),     # beyond the warp kernel's n <= 2048: 4-CTA clusters
    (3000, 21, 40, 0.7, False),    # 6-CTA clusters (the C3 sample size)
    (5000, 17, 24, 1.0, True),     # 8-CTA clusters (the C5 sample size)
])
def test_single_sweep_parity_missing(oracle_built, mis_kernel, n, p, q, c, shuffle):
    from atlasqtl_b200.device import SweepContext
    native = oracle_built
    if mis_kernel == "warp" and n > 2048:
        pytest.skip("the warp-per-trait kernel covers n <= 2048")
    X, Y, hyper, init = make_problem(n, p, q)
    p, q = X.shape[1], Y.shape[1]
    si = sweep_inputs(X, Y, init, c=c)
    mi = mis_inputs(X, Y, si)
    order = (np.random.default_rng(5).permutation(p) if shuffle else np.arange(p)).astype(np.int32)
    gam, mu = si["gam"].copy(order="F"), si["mu"].copy(order="F")
    beta = np.asfortranarray(gam * mu)
    R0 = np.asfortranarray(mi["mis"] * (mi["Y"] - X @ beta))
    R = R0.copy(order="F")
    native.sweep_primal_mis(X, mi["mis"], mi["xnsq"], R, gam, si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"],
                            si["log_tau"], beta, mu, mi["sig2_beta"], si["tau"], order, c=c)
    Ynan = np.where(mi["mis"] == 0, np.nan, Y)
    with SweepContext(X, np.where(np.isnan(Ynan), 123.0, Ynan)) as ctx:   # whatever sits in the missing slots is ignored
        n_obs = ctx.set_missing(mi["mis"])
        np.testing.assert_array_equal(n_obs, mi["mis"].sum(axis=0))
        ctx.set_order(order)
        st0 = ctx.set_state_mis(si["gam"], si["mu"])
        b0 = si["gam"] * si["mu"]
        np.testing.assert_allclose(ctx.get_residual(), R0, atol=1e-10)
        np.testing.assert_allclose(st0["resid_sq"], (R0 ** 2).sum(axis=0), rtol=1e-11)
        np.testing.assert_allclose(st0["colsum_xn_gam"], (mi["xnsq"] * si["gam"]).sum(axis=0), rtol=1e-11)
        np.testing.assert_allclose(st0["colsum_xn_beta2"], (mi["xnsq"] * b0 ** 2).sum(axis=0), rtol=1e-11)
        np.testing.assert_allclose(st0["colsum_xn_gam_mu2"], (mi["xnsq"] * si["gam"] * si["mu"] ** 2).sum(axis=0), rtol=1e-11)
        np.testing.assert_allclose(st0["colsum_beta2"], (b0 ** 2).sum(axis=0), rtol=1e-11)
        ctx.refresh_tables(si["theta"], si["zeta"], c_next=c)
        out = ctx.sweep_mis(c, si["log_sig2_inv"], mi["sig2_inv"], si["tau"], si["log_tau"])
        st = ctx.get_state()
        Rg = ctx.get_residual()
        with pytest.raises(Exception, match="use aq_sweep_mis"):
            ctx.sweep(c, 0.0, si["tau"], si["log_tau"], si["tau"])
    assert np.abs(st["gam_vb"] - gam).max() <= 1e-9      # north_star bound: 1e-8
    assert np.abs(st["mu_beta_vb"] - mu).max() <= 1e-9
    assert np.abs(st["beta_vb"] - beta).max() <= 1e-9
    np.testing.assert_allclose(Rg, R, atol=1e-9)
    assert np.all(Rg[mi["mis"] == 0] == 0.0)               # the residual stays masked
    s2 = mi["sig2_beta"]
    np.testing.assert_allclose(out["colsum_gam"], gam.sum(axis=0), rtol=1e-9)
    np.testing.assert_allclose(out["colsum_gam_mu2"] + out["colsum_sig2b_gam"], ((mu ** 2 + s2) * gam).sum(axis=0), rtol=1e-9)
    np.testing.assert_allclose(out["colsum_xn_gam_mu2"] + out["colsum_xn_sig2b_gam"],
                               (mi["xnsq"] * (mu ** 2 + s2) * gam).sum(axis=0), rtol=1e-9)
    np.testing.assert_allclose(out["colsum_xn_beta2"], (mi["xnsq"] * beta ** 2).sum(axis=0), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out["colsum_gam_logsig2b"], (gam * np.log(s2)).sum(axis=0), rtol=1e-9)
    np.testing.assert_allclose(out["resid_sq"], (R ** 2).sum(axis=0), rtol=1e-9)

    if p <= 40 and native.ref_available():   # the reference's own loop (p x p x q inputs)
        cp_X = np.asfortranarray(X.T @ X)
        cp_X_rm = np.zeros((p, p, q), order="F")
        for k in range(q):
            rows = np.flatnonzero(mi["mis"][:, k] == 0)
            cp_X_rm[:, :, k] = X[rows].T @ X[rows]
        g_r, m_r = si["gam"].copy(order="F"), si["mu"].copy(order="F")
        b_r = np.asfortranarray(g_r * m_r)
        cbx = np.asfortranarray(cp_X @ b_r - np.stack([cp_X_rm[:, :, k] @ b_r[:, k] for k in range(q)], axis=1))
        native.ref_core_dual_mis_loop(cp_X, cp_X_rm, np.asfortranarray(mi["Y"].T @ X), g_r, si["log_Phi"],
                                      si["log_1_min_Phi"], si["log_sig2_inv"], si["log_tau"], b_r, cbx, m_r, s2, si["tau"],
                                      order, np.arange(q, dtype=np.int32), c=c)
        assert np.abs(st["gam_vb"] - g_r).max() <= 1e-9
        assert np.abs(st["beta_vb"] - b_r).max() <= 1e-9
        np.testing.assert_allclose(X.T @ (mi["Y"] - Rg), cbx, atol=1e-8)


@pytest.mark.parametrize("n,p,q,anneal", [(100, 75, 20, (1, 2, 5)), (300, 120, 40, None)])
def test_trajectory_parity_missing(oracle_built, n, p, q, anneal):
    """Full VB runs with NaN responses: same iteration count, ELBO trajectory and selections as the restated R loop."""
    from atlasqtl_b200 import core
    from oracle import vb_oracle
    X, Y, hyper, init = make_problem(n, p, q)
    q = Y.shape[1]
    Ym = Y.copy()
    Ym[np.random.default_rng(3).uniform(size=Y.shape) < 0.06] = np.nan
    tr_o, tr_g = [], []
    ref = vb_oracle.atlasqtl_global_local_core_(Ym, X, q, anneal, 1, 0.1, 1000, hyper, init, trace=tr_o)
    out = core.atlasqtl_global_local_core_(Ym, X, q, anneal, 1, 0.1, 1000, 0, hyper, init, debug=True, trace=tr_g)
    assert out["converged"] and ref["converged"]
    assert out["it"] == ref["it"]
    for a, b in zip(tr_o, tr_g):
        if a["lb"] is not None:
            assert abs(a["lb"] - b["lb"]) <= 1e-10 * abs(a["lb"]), (a["it"], a["lb"], b["lb"])
    assert np.abs(out["gam_vb"] - ref["gam_vb"]).max() <= 1e-8
    assert np.abs(out["beta_vb"] - ref["beta_vb"]).max() <= 1e-8
    assert np.array_equal(out["gam_vb"] > 0.5, ref["gam_vb"] > 0.5)


def test_all_observed_pattern_reproduces_the_fast_path(oracle_built):
    """mis_pat == 1 everywhere: coreDualMisLoop degenerates to coreDualLoop with sig2_beta_vb(j,k) = sig2_beta_vb[k]
    (X_norm_sq = n - 1 for standardised X), so the two CUDA paths must agree."""
    from atlasqtl_b200.device import SweepContext
    X, Y, hyper, init = make_problem(300, 90, 37)
    p, q = X.shape[1], Y.shape[1]
    n = X.shape[0]
    c = 0.8
    si = sweep_inputs(X, Y, init, c=c)
    sig2_inv = 0.7   # sweep_inputs builds sig2_beta = 1 / (c (n - 1 + 0.7) tau)
    order = np.random.default_rng(2).permutation(p).astype(np.int32)
    with SweepContext(X, Y) as ctx:
        ctx.set_order(order)
        ctx.set_state(si["gam"], si["mu"])
        ctx.refresh_tables(si["theta"], si["zeta"], c_next=c)
        a = ctx.sweep(c, si["log_sig2_inv"], si["tau"], si["log_tau"], si["sig2_beta"])
        sa = ctx.get_state()
    with SweepContext(X, Y) as ctx:
        n_obs = ctx.set_missing(np.ones((n, q)))
        assert np.all(n_obs == n)
        ctx.set_order(order)
        ctx.set_state_mis(si["gam"], si["mu"])
        ctx.refresh_tables(si["theta"], si["zeta"], c_next=c)
        b = ctx.sweep_mis(c, si["log_sig2_inv"], sig2_inv, si["tau"], si["log_tau"])
        sb = ctx.get_state()
    assert np.abs(sa["gam_vb"] - sb["gam_vb"]).max() <= 1e-9
    assert np.abs(sa["mu_beta_vb"] - sb["mu_beta_vb"]).max() <= 1e-9
    np.testing.assert_allclose(a["colsum_gam"], b["colsum_gam"], rtol=1e-9)
    np.testing.assert_allclose(a["resid_sq"], b["resid_sq"], rtol=1e-9)
    np.testing.assert_allclose(a["colsum_gam_mu2"] + si["sig2_beta"] * a["colsum_gam"],
                               b["colsum_gam_mu2"] + b["colsum_sig2b_gam"], rtol=1e-8)
    np.testing.assert_allclose(a["colsum_zpart"], b["colsum_zpart"], rtol=1e-9, atol=1e-9)
