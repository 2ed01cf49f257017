"""ctypes loaders for the oracle's native libraries -- TEST INFRASTRUCTURE ONLY.

``liboracle.so``       built from oracle/cavi_oracle.c (our restatement)
``libatlasqtl_ref.so`` the reference's own src/coreLoop.cpp compiled unmodified (oracle/_ref/)

All matrices are Fortran-ordered float64 (R layout); index vectors int32, 0-based
(R/atlasqtl_global_local_core.R:162-163).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DP = ctypes.POINTER(ctypes.c_double)
_IP = ctypes.POINTER(ctypes.c_int)


def build(verbose=False):
    """Run oracle/Makefile (gcc/g++ only).  _ref is rebuilt only where /root/reference exists."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def _load(path):
    if not os.path.exists(path):
        build()
    return ctypes.CDLL(path)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load(os.path.join(_HERE, "_build", "liboracle.so"))
    return _lib


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libatlasqtl_ref.so")) or os.path.exists(
        "/root/reference/src/coreLoop.cpp")


def ref():
    global _ref
    if _ref is None:
        _ref = _load(os.path.join(_HERE, "_ref", "libatlasqtl_ref.so"))
    return _ref


def _d(a):
    assert a.dtype == np.float64 and a.flags.f_contiguous, "need Fortran-ordered float64"
    return a.ctypes.data_as(_DP)


def _i(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_IP)


def core_dual_loop(cp_X, cp_Y_X, gam_vb, log_Phi, log_1_min_Phi, log_sig2_inv_vb, log_tau_vb,
                   m1_beta, cp_betaX_X, mu_beta_vb, sig2_beta_vb, tau_vb, shuffled_ind, sample_q,
                   c=1.0, impl="oracle"):
    """Same argument order as the R closure (R/RcppExports.R:4-6).  In place on
    gam_vb, m1_beta, cp_betaX_X, mu_beta_vb.  impl: "oracle" (cavi_oracle.c) or
    "reference" (the reference's own coreLoop.cpp)."""
    p, q = gam_vb.shape
    fn = lib().oracle_core_dual_loop if impl == "oracle" else ref().ref_coreDualLoop
    fn.restype = None
    fn(ctypes.c_int(p), ctypes.c_int(q), _d(cp_X), _d(cp_Y_X), _d(gam_vb), _d(log_Phi),
       _d(log_1_min_Phi), ctypes.c_double(log_sig2_inv_vb), _d(log_tau_vb), _d(m1_beta),
       _d(cp_betaX_X), _d(mu_beta_vb), _d(sig2_beta_vb), _d(tau_vb), _i(shuffled_ind),
       ctypes.c_int(len(shuffled_ind)), _i(sample_q), ctypes.c_int(len(sample_q)),
       ctypes.c_double(c))


def ref_core_dual_mis_loop(cp_X, cp_X_rm, cp_Y_X, gam_vb, log_Phi, log_1_min_Phi, log_sig2_inv_vb,
                           log_tau_vb, m1_beta, cp_betaX_X, mu_beta_vb, sig2_beta_vb, tau_vb,
                           shuffled_ind, sample_q, c=1.0):
    """The reference's coreDualMisLoop.  cp_X_rm: (p, p, q) Fortran array = q stacked p x p matrices."""
    p, q = gam_vb.shape
    fn = ref().ref_coreDualMisLoop
    fn.restype = None
    fn(ctypes.c_int(p), ctypes.c_int(q), _d(cp_X), _d(cp_X_rm), _d(cp_Y_X), _d(gam_vb), _d(log_Phi),
       _d(log_1_min_Phi), ctypes.c_double(log_sig2_inv_vb), _d(log_tau_vb), _d(m1_beta),
       _d(cp_betaX_X), _d(mu_beta_vb), _d(sig2_beta_vb), _d(tau_vb), _i(shuffled_ind),
       ctypes.c_int(len(shuffled_ind)), _i(sample_q), ctypes.c_int(len(sample_q)),
       ctypes.c_double(c))


def sweep_primal(X, xnorm2, R, gam_vb, log_Phi, log_1_min_Phi, log_sig2_inv_vb, log_tau_vb, m1_beta,
                 mu_beta_vb, sig2_beta_vb, tau_vb, shuffled_ind, c=1.0, nthreads=1):
    n, p = X.shape
    q = R.shape[1]
    fn = lib().oracle_sweep_primal
    fn.restype = None
    fn(ctypes.c_int(n), ctypes.c_int(p), ctypes.c_int(q), _d(X), _d(xnorm2), _d(R), _d(gam_vb),
       _d(log_Phi), _d(log_1_min_Phi), ctypes.c_double(log_sig2_inv_vb), _d(log_tau_vb), _d(m1_beta),
       _d(mu_beta_vb), _d(sig2_beta_vb), _d(tau_vb), _i(shuffled_ind), ctypes.c_int(len(shuffled_ind)),
       ctypes.c_double(c), ctypes.c_int(nthreads))


def sweep_primal_blocked(X, R, gam_vb, log_Phi, log_1_min_Phi, log_sig2_inv_vb, log_tau_vb, m1_beta,
                         mu_beta_vb, sig2_beta_vb, tau_vb, shuffled_ind, c=1.0, B=8):
    n, p = X.shape
    q = R.shape[1]
    fn = lib().oracle_sweep_primal_blocked
    fn.restype = None
    fn(ctypes.c_int(n), ctypes.c_int(p), ctypes.c_int(q), ctypes.c_int(B), _d(X), _d(R), _d(gam_vb),
       _d(log_Phi), _d(log_1_min_Phi), ctypes.c_double(log_sig2_inv_vb), _d(log_tau_vb), _d(m1_beta),
       _d(mu_beta_vb), _d(sig2_beta_vb), _d(tau_vb), _i(shuffled_ind), ctypes.c_int(len(shuffled_ind)),
       ctypes.c_double(c))


def sweep_primal_mis(X, mis, xnsq, R, gam_vb, log_Phi, log_1_min_Phi, log_sig2_inv_vb, log_tau_vb, m1_beta,
                     mu_beta_vb, sig2_beta_vb, tau_vb, shuffled_ind, c=1.0):
    """coreDualMisLoop in sample space: mis n x q (1 observed / 0 missing), xnsq = crossprod(X^2, mis) and
    sig2_beta_vb p x q, R = mis * (Y - X beta) in / out."""
    n, p = X.shape
    q = R.shape[1]
    fn = lib().oracle_sweep_primal_mis
    fn.restype = None
    fn(ctypes.c_int(n), ctypes.c_int(p), ctypes.c_int(q), _d(X), _d(mis), _d(xnsq), _d(R), _d(gam_vb),
       _d(log_Phi), _d(log_1_min_Phi), ctypes.c_double(log_sig2_inv_vb), _d(log_tau_vb), _d(m1_beta),
       _d(mu_beta_vb), _d(sig2_beta_vb), _d(tau_vb), _i(shuffled_ind), ctypes.c_int(len(shuffled_ind)),
       ctypes.c_double(c))


def residual(X, Y, beta):
    n, p = X.shape
    q = Y.shape[1]
    R = np.empty((n, q), order="F")
    fn = lib().oracle_residual
    fn.restype = None
    fn(ctypes.c_int(n), ctypes.c_int(p), ctypes.c_int(q), _d(X), _d(Y), _d(beta), _d(R))
    return R
