"""GPU parity of one sweep through the C ABI against the CPU oracle (and, at dual-feasible sizes,
against the reference's own coreLoop.cpp).  Tolerances: max|d gam_vb| <= 1e-8 (north_star), tighter in practice."""
import numpy as np
import pytest

from problems import make_problem, sweep_inputs

pytestmark = pytest.mark.gpu

LOG_SQRT_2PI = 0.5 * np.log(2 * np.pi)


def oracle_sweep(native, X, Y, si, order, form):
    gam, mu = si["gam"].copy(order="F"), si["mu"].copy(order="F")
    beta = np.asfortranarray(gam * mu)
    q = Y.shape[1]
    if form == "reference":
        cp_X = np.asfortranarray(X.T @ X)
        cp_Y_X = np.asfortranarray(Y.T @ X)
        cbx = np.asfortranarray(cp_X @ beta)
        native.core_dual_loop(cp_X, cp_Y_X, gam, si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"], si["log_tau"],
                              beta, cbx, mu, si["sig2_beta"], si["tau"], order, np.arange(q, dtype=np.int32),
                              c=si["c"], impl="reference")
        R = np.asfortranarray(Y - X @ beta)
    else:
        R = native.residual(X, Y, beta)
        xn = np.asfortranarray(np.sum(X ** 2, axis=0))
        native.sweep_primal(X, xn, R, gam, si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"], si["log_tau"], beta,
                            mu, si["sig2_beta"], si["tau"], order, c=si["c"], nthreads=4)
    return gam, mu, beta, R


def zparts(si, gam):
    from scipy import special as sp
    u = si["theta"][:, None] + si["zeta"][None, :]
    sc = 1.0 if abs(si["c"] - 1) < 1.5e-8 else np.sqrt(si["c"])
    U = sc * u
    lp, lq = sp.log_ndtr(U), sp.log_ndtr(-U)
    m1 = np.exp(-U ** 2 / 2 - LOG_SQRT_2PI - lp)
    m1 = np.where(m1 < -U, -U, m1)
    m0 = -np.exp(-U ** 2 / 2 - LOG_SQRT_2PI - lq)
    m0 = np.where(m0 > -U, -U, m0)
    zp = gam * (m1 - m0) + m0
    return zp.sum(axis=0), zp.sum(axis=1)


@pytest.mark.parametrize("n,p,q,c,form,shuffle", [
    (100, 75, 20, 1.0, "reference", False),
    (100, 75, 20, 0.6, "reference", True),
    (200, 300, 200, 0.8, "reference", True),
    (500, 203, 70, 1.0, "primal", True),
    (1000, 120, 50, 0.5, "primal", False),
    # one full round of the persistent grid (148 tiles) + a tail launch: the row sums come from the per-tile partials
    # the helper warps leave (main launch) plus one streamed row (tail traits)
    (1000, 21, 2500, 0.5, "primal", True),    # 16-trait tiles, 17 tail tiles of 8
    (100, 24, 4776, 0.7, "primal", True),     # 32-trait tiles (one trait per helper lane), 5 tail tiles
    (600, 19, 3552 + 3, 1.0, "primal", False),  # 24-trait tiles, q not a multiple of 8
])
def test_single_sweep_parity(oracle_built, n, p, q, c, form, shuffle):
    from atlasqtl_b200.device import SweepContext
    native = oracle_built
    X, Y, hyper, init = make_problem(n, p, q)
    p = X.shape[1]
    si = sweep_inputs(X, Y, init, c=c)
    order = (np.random.default_rng(5).permutation(p) if shuffle else np.arange(p)).astype(np.int32)
    g_ref, m_ref, b_ref, R_ref = oracle_sweep(native, X, Y, si, order, form)

    with SweepContext(X, Y) as ctx:
        ctx.set_order(order)
        st0 = ctx.set_state(si["gam"], si["mu"])
        beta0 = si["gam"] * si["mu"]
        np.testing.assert_allclose(st0["colsum_gam"], si["gam"].sum(axis=0), rtol=1e-12)
        np.testing.assert_allclose(st0["colsum_beta2"], (beta0 ** 2).sum(axis=0), rtol=1e-12)
        np.testing.assert_allclose(st0["colsum_gam_mu2"], (si["gam"] * si["mu"] ** 2).sum(axis=0), rtol=1e-12)
        R0 = Y - X @ beta0
        np.testing.assert_allclose(ctx.get_residual(), R0, atol=1e-10)
        np.testing.assert_allclose(st0["resid_sq"], (R0 ** 2).sum(axis=0), rtol=1e-11)

        ctx.refresh_tables(si["theta"], si["zeta"], c_next=c)
        out = ctx.sweep(c, si["log_sig2_inv"], si["tau"], si["log_tau"], si["sig2_beta"])
        st = ctx.get_state()
        R = ctx.get_residual()
        rows = ctx.rowsums_zpart()

    assert np.abs(st["gam_vb"] - g_ref).max() <= 1e-9
    assert np.abs(st["mu_beta_vb"] - m_ref).max() <= 1e-9 * max(1.0, np.abs(m_ref).max())
    assert np.abs(st["beta_vb"] - b_ref).max() <= 1e-9
    np.testing.assert_allclose(R, R_ref, atol=1e-9)
    np.testing.assert_allclose(out["colsum_gam"], g_ref.sum(axis=0), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(out["colsum_beta2"], (b_ref ** 2).sum(axis=0), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(out["colsum_gam_mu2"], (g_ref * m_ref ** 2).sum(axis=0), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(out["resid_sq"], (R_ref ** 2).sum(axis=0), rtol=1e-9)
    zc, zr = zparts(si, g_ref)
    np.testing.assert_allclose(out["colsum_zpart"], zc, rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(rows, zr, rtol=1e-9, atol=1e-8)
