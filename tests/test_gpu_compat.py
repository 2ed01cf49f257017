"""GPU: the stateless drop-in `coreDualLoop` (same 15 arguments, same in-place outputs as the reference's .Call
entry point) against the reference's own compiled coreDualLoop on identical arguments."""
import numpy as np
import pytest

from problems import make_problem, sweep_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,p,q,c,subset", [(100, 75, 20, 1.0, False), (80, 130, 12, 0.7, True), (200, 300, 40, 0.9, False)])
def test_coredualloop_dropin(oracle_built, n, p, q, c, subset):
    from atlasqtl_b200.compat import coreDualLoop
    native = oracle_built
    X, Y, hyper, init = make_problem(n, p, q)
    p = X.shape[1]
    si = sweep_inputs(X, Y, init, c=c)
    order = np.random.default_rng(9).permutation(p).astype(np.int32)
    sample_q = (np.array([5, 1, 8], dtype=np.int32) if subset else np.arange(q, dtype=np.int32))
    cp_X, cp_Y_X = np.asfortranarray(X.T @ X), np.asfortranarray(Y.T @ X)

    def fresh():
        gam, mu = si["gam"].copy(order="F"), si["mu"].copy(order="F")
        beta = np.asfortranarray(gam * mu)
        return gam, mu, beta, np.asfortranarray(cp_X @ beta)

    g_r, m_r, b_r, cbx_r = fresh()
    native.core_dual_loop(cp_X, cp_Y_X, g_r, si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"], si["log_tau"], b_r,
                          cbx_r, m_r, si["sig2_beta"], si["tau"], order, sample_q, c=c,
                          impl="reference" if native.ref_available() else "oracle")
    g, m, b, cbx = fresh()
    coreDualLoop(cp_X, cp_Y_X, g, si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"], si["log_tau"], b, cbx, m,
                 si["sig2_beta"], si["tau"], order, sample_q, c=c)
    assert np.abs(g - g_r).max() <= 1e-8
    assert np.abs(m - m_r).max() <= 1e-8
    assert np.abs(b - b_r).max() <= 1e-8
    np.testing.assert_allclose(cbx, cbx_r, atol=1e-7)
    if subset:  # untouched columns stay bit-identical, like the reference
        rest = np.setdiff1d(np.arange(q), sample_q)
        assert np.array_equal(g[:, rest], si["gam"][:, rest])


@pytest.mark.parametrize("n,p,q,c,subset", [(60, 40, 12, 0.7, False), (100, 33, 9, 1.0, True), (400, 56, 7, 0.9, False)])
def test_coredualmisloop_dropin(oracle_built, n, p, q, c, subset):
    """The stateless 16-argument `coreDualMisLoop` (src/RcppExports.cpp:41-63) against the reference's own compiled loop on
    identical arguments: cp_X_rm = list of per-trait crossprod(X[missing rows, ]) (R/atlasqtl_global_local_core.R:25-32),
    an ARBITRARY p x q sig2_beta_vb (the entry takes it as an argument; it need not be update_sig2_beta_vb_'s value)."""
    from atlasqtl_b200.compat import coreDualMisLoop
    from problems import mis_inputs
    native = oracle_built
    if not native.ref_available():
        pytest.skip("oracle/_ref (the reference's coreLoop.cpp) is not available on this box")
    X, Y, hyper, init = make_problem(n, p, q)
    p, q = X.shape[1], Y.shape[1]
    si = sweep_inputs(X, Y, init, c=c)
    mi = mis_inputs(X, Y, si)
    rng = np.random.default_rng(4)
    sig2 = np.asfortranarray(mi["sig2_beta"] * rng.uniform(0.5, 2.0, size=(p, q)))
    order = rng.permutation(p).astype(np.int32)
    sample_q = (np.array([5, 1, 8], dtype=np.int32) if subset else np.arange(q, dtype=np.int32))
    cp_X = np.asfortranarray(X.T @ X)
    cp_Y_X = np.asfortranarray(mi["Y"].T @ X)
    rm = [np.asfortranarray(X[mi["mis"][:, k] == 0].T @ X[mi["mis"][:, k] == 0]) for k in range(q)]
    rm3 = np.asfortranarray(np.stack(rm, axis=2))

    def fresh():
        gam, mu = si["gam"].copy(order="F"), si["mu"].copy(order="F")
        beta = np.asfortranarray(gam * mu)
        cbx = np.asfortranarray(np.stack([(cp_X - rm[k]) @ beta[:, k] for k in range(q)], axis=1))
        return gam, mu, beta, cbx

    g_r, m_r, b_r, cbx_r = fresh()
    native.ref_core_dual_mis_loop(cp_X, rm3, cp_Y_X, g_r, si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"],
                                  si["log_tau"], b_r, cbx_r, m_r, sig2, si["tau"], order, sample_q, c=c)
    g, m, b, cbx = fresh()
    coreDualMisLoop(cp_X, rm, cp_Y_X, g, si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"], si["log_tau"], b, cbx, m,
                    sig2, si["tau"], order, sample_q, c=c)
    assert np.abs(g - g_r).max() <= 1e-8
    assert np.abs(m - m_r).max() <= 1e-8 * max(1.0, np.abs(m_r).max())
    assert np.abs(b - b_r).max() <= 1e-8
    np.testing.assert_allclose(cbx, cbx_r, atol=1e-7)
    if subset:
        rest = np.setdiff1d(np.arange(q), sample_q)
        assert np.array_equal(g[:, rest], si["gam"][:, rest])
        assert np.array_equal(cbx[:, rest], fresh()[3][:, rest])
