"""Seeded test problems shared by the CPU and GPU tests (inputs only; no expected values here)."""
import numpy as np
from scipy import special as sp

from atlasqtl_b200 import hyper_init, synthetic


def make_problem(n, p, q, p_act=None, q_act=None, seed=123, p0=None, maf=0.25):
    p_act = p_act or max(2, p // 10)
    q_act = q_act or max(2, q // 2)
    X, Y, pat = synthetic.simulate(n, p, q, p_act, q_act, maf=maf, seed=seed)
    p = X.shape[1]
    p0 = p0 or (max(1.0, float(pat.sum(axis=0).mean())), 10.0)
    hyper = hyper_init.auto_set_hyper_(Y, p, p0)
    init = hyper_init.auto_set_init_(Y, p, p0, q, user_seed=seed)
    return X, Y, hyper, init


def sweep_inputs(X, Y, init, c=1.0, seed=7):
    """A self-consistent set of single-sweep inputs (what R/atlasqtl_global_local_core.R:134-150 would hand
    to coreDualLoop), randomised so that every term of the update matters."""
    rng = np.random.default_rng(seed)
    n, p = X.shape
    q = Y.shape[1]
    gam = np.asfortranarray(rng.uniform(0.0, 1.0, size=(p, q)) ** 4)
    mu = np.asfortranarray(rng.normal(0.0, 0.3, size=(p, q)))
    theta = rng.normal(0.0, 0.5, size=p)
    zeta = rng.normal(-1.5, 0.5, size=q)
    tau = rng.uniform(0.5, 2.0, size=q)
    sig2_inv = 0.7
    sig2_beta = 1.0 / (c * (n - 1 + sig2_inv) * tau)
    log_tau = np.log(tau) - 0.01
    log_sig2_inv = float(np.log(sig2_inv) - 0.02)
    u = theta[:, None] + zeta[None, :]
    log_Phi = np.asfortranarray(sp.log_ndtr(u))
    log_1_min_Phi = np.asfortranarray(sp.log_ndtr(-u))
    return dict(gam=gam, mu=mu, theta=theta, zeta=zeta, tau=tau, sig2_beta=sig2_beta, log_tau=log_tau,
                log_sig2_inv=log_sig2_inv, log_Phi=log_Phi, log_1_min_Phi=log_1_min_Phi, c=c)


def mis_inputs(X, Y, si, frac=0.08, seed=11):
    """Missing-response version of `sweep_inputs` (what R/atlasqtl_global_local_core.R:19-33, :147 hand to
    coreDualMisLoop): mis n x q (1 observed / 0 missing; some traits fully observed), Y zeroed where missing,
    X_norm_sq = crossprod(X^2, mis) and the p x q sig2_beta_vb of update_sig2_beta_vb_ (R/update_vb.R:47)."""
    rng = np.random.default_rng(seed)
    n, q = Y.shape
    mis = (rng.uniform(size=(n, q)) >= frac).astype(np.float64)
    mis[:, ::3] = 1.0
    mis = np.asfortranarray(mis)
    Ym = np.asfortranarray(Y * mis)
    xnsq = np.asfortranarray((X ** 2).T @ mis)
    sig2_inv = 0.7
    sig2_beta = np.asfortranarray(1.0 / (si["c"] * (xnsq + sig2_inv) * si["tau"][None, :]))
    return dict(mis=mis, Y=Ym, xnsq=xnsq, sig2_inv=sig2_inv, sig2_beta=sig2_beta)
