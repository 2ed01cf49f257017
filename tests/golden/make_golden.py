"""Generates tests/golden/coredualloop_*.npz by running the REFERENCE's own coreDualLoop / coreDualMisLoop
(src/coreLoop.cpp compiled unmodified into oracle/_ref, see oracle/Makefile) on seeded inputs.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The fixtures travel with the repo, so the GPU box (no /root/reference) checks against the same vectors.
"""
import os
import sys

import numpy as np
from scipy import special as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import native  # noqa: E402

CASES = {  # name: (n, p, q, c, shuffled, seed)
    "a_identity_c1": (60, 40, 12, 1.0, False, 1),
    "b_shuffled_c07": (80, 57, 9, 0.7, True, 2),
    "c_wide_c05": (50, 130, 7, 0.5, True, 3),
}


def inputs(n, p, q, c, shuffled, seed):
    rng = np.random.default_rng(seed)
    G = rng.binomial(2, 0.3, size=(n, p)).astype(np.float64)
    G[:2, G.std(axis=0) == 0] = [[0.0], [1.0]]
    X = np.asfortranarray((G - G.mean(0)) / G.std(0, ddof=1))
    Y = rng.normal(size=(n, q)) + X[:, :3] @ rng.normal(size=(3, q))
    Y = np.asfortranarray(Y - Y.mean(0))
    gam = np.asfortranarray(rng.uniform(size=(p, q)) ** 3)
    mu = np.asfortranarray(rng.normal(0, 0.4, size=(p, q)))
    theta, zeta = rng.normal(0, 0.6, p), rng.normal(-1.2, 0.4, q)
    tau = rng.uniform(0.5, 2.0, q)
    sig2 = 1.0 / (c * (n - 1 + 0.8) * tau)
    order = (rng.permutation(p) if shuffled else np.arange(p)).astype(np.int32)
    return dict(X=X, Y=Y, gam=gam, mu=mu, theta=theta, zeta=zeta, tau=tau, sig2_beta=sig2, log_tau=np.log(tau) - 0.03,
                log_sig2_inv=-0.4, c=c, order=order)


def run_reference(d):
    X, Y = d["X"], d["Y"]
    q = Y.shape[1]
    u = d["theta"][:, None] + d["zeta"][None, :]
    lp, lq = np.asfortranarray(sp.log_ndtr(u)), np.asfortranarray(sp.log_ndtr(-u))
    gam, mu = d["gam"].copy(order="F"), d["mu"].copy(order="F")
    beta = np.asfortranarray(gam * mu)
    cp_X, cp_Y_X = np.asfortranarray(X.T @ X), np.asfortranarray(Y.T @ X)
    cbx = np.asfortranarray(cp_X @ beta)
    native.core_dual_loop(cp_X, cp_Y_X, gam, lp, lq, d["log_sig2_inv"], d["log_tau"], beta, cbx, mu, d["sig2_beta"],
                          d["tau"], d["order"], np.arange(q, dtype=np.int32), c=d["c"], impl="reference")
    return dict(log_Phi=lp, log_1_min_Phi=lq, out_gam=gam, out_mu=mu, out_beta=beta, out_cp_betaX_X=cbx)


def run_reference_mis(d, frac=0.1, seed=17):
    """The reference's coreDualMisLoop (src/coreLoop.cpp:91-138) with the set-up of R/atlasqtl_global_local_core.R:19-33:
    mis_pat, Y zeroed where missing, X_norm_sq, the list cp_X_rm of per-trait Gram corrections, p x q sig2_beta_vb."""
    X, Y = d["X"], d["Y"]
    n, p = X.shape
    q = Y.shape[1]
    rng = np.random.default_rng(seed)
    mis = (rng.uniform(size=(n, q)) >= frac).astype(np.float64)
    mis[:, ::3] = 1.0
    mis = np.asfortranarray(mis)
    Ym = np.asfortranarray(Y * mis)
    xnsq = np.asfortranarray((X ** 2).T @ mis)
    sig2_inv = 0.8
    sig2_beta = np.asfortranarray(1.0 / (d["c"] * (xnsq + sig2_inv) * d["tau"][None, :]))
    u = d["theta"][:, None] + d["zeta"][None, :]
    lp, lq = np.asfortranarray(sp.log_ndtr(u)), np.asfortranarray(sp.log_ndtr(-u))
    cp_X = np.asfortranarray(X.T @ X)
    cp_X_rm = np.zeros((p, p, q), order="F")
    for k in range(q):
        rows = np.flatnonzero(mis[:, k] == 0)
        cp_X_rm[:, :, k] = X[rows].T @ X[rows]
    cp_Y_X = np.asfortranarray(Ym.T @ X)
    gam, mu = d["gam"].copy(order="F"), d["mu"].copy(order="F")
    beta = np.asfortranarray(gam * mu)
    cbx = np.asfortranarray(cp_X @ beta - np.stack([cp_X_rm[:, :, k] @ beta[:, k] for k in range(q)], axis=1))
    native.ref_core_dual_mis_loop(cp_X, cp_X_rm, cp_Y_X, gam, lp, lq, d["log_sig2_inv"], d["log_tau"], beta, cbx, mu,
                                  sig2_beta, d["tau"], d["order"], np.arange(q, dtype=np.int32), c=d["c"])
    return dict(mis=mis, Y_mis=Ym, xnsq=xnsq, sig2_inv=sig2_inv, sig2_beta_pq=sig2_beta, log_Phi=lp, log_1_min_Phi=lq,
                out_gam=gam, out_mu=mu, out_beta=beta, out_cp_betaX_X=cbx)


if __name__ == "__main__":
    native.build()
    for name, spec in (("a_shuffled_c07", (80, 57, 9, 0.7, True, 2)), ("b_identity_c1", (70, 33, 11, 1.0, False, 5))):
        d = inputs(*spec)
        d.update(run_reference_mis(d))
        np.savez_compressed(os.path.join(HERE, f"coredualmisloop_{name}.npz"), **d)
        print("mis", name, {k: getattr(v, "shape", v) for k, v in d.items() if k.startswith("out_")})
    for name, spec in CASES.items():
        d = inputs(*spec)
        d.update(run_reference(d))
        np.savez_compressed(os.path.join(HERE, f"coredualloop_{name}.npz"), **d)
        print(name, {k: getattr(v, "shape", v) for k, v in d.items() if k.startswith("out_")})
