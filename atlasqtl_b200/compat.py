"""`coreDualLoop` / `coreDualMisLoop`: the reference's R closures (R/RcppExports.R:4-10) with the same 15 / 16 arguments
and the same in-place outputs, backed by the CUDA sweeps (aq_coreDualLoop / aq_coreDualMisLoop, include/atlasqtl_b200.h)."""
import ctypes

import numpy as np

from . import _lib


def coreDualLoop(cp_X, cp_Y_X, gam_vb, log_Phi_theta_plus_zeta, log_1_min_Phi_theta_plus_zeta, log_sig2_inv_vb,
                 log_tau_vb, m1_beta, cp_betaX_X, mu_beta_vb, sig2_beta_vb, tau_vb, shuffled_ind, sample_q, c=1.0,
                 device=0):
    """In place on gam_vb, m1_beta, cp_betaX_X, mu_beta_vb (Fortran-ordered float64), like the reference."""
    lib = _lib.load()
    p, q = gam_vb.shape
    mats = (cp_X, cp_Y_X, gam_vb, log_Phi_theta_plus_zeta, log_1_min_Phi_theta_plus_zeta, m1_beta, cp_betaX_X, mu_beta_vb)
    for m in mats:
        if not (isinstance(m, np.ndarray) and m.dtype == np.float64 and m.flags.f_contiguous):
            raise TypeError("matrices must be Fortran-ordered float64 arrays (R's double storage)")
    if cp_X.shape != (p, p) or cp_Y_X.shape != (q, p):
        raise ValueError("cp_X must be p x p and cp_Y_X q x p")
    vecs = [np.ascontiguousarray(v, dtype=np.float64) for v in (log_tau_vb, sig2_beta_vb, tau_vb)]
    si = np.ascontiguousarray(shuffled_ind, dtype=np.int32)
    sq = np.ascontiguousarray(sample_q, dtype=np.int32)
    _lib.check(lib.aq_coreDualLoop(ctypes.c_int(device), p, q, _lib.dptr(cp_X), _lib.dptr(cp_Y_X), _lib.dptr(gam_vb),
                                   _lib.dptr(log_Phi_theta_plus_zeta), _lib.dptr(log_1_min_Phi_theta_plus_zeta),
                                   ctypes.c_double(log_sig2_inv_vb), _lib.dptr(vecs[0]), _lib.dptr(m1_beta),
                                   _lib.dptr(cp_betaX_X), _lib.dptr(mu_beta_vb), _lib.dptr(vecs[1]), _lib.dptr(vecs[2]),
                                   _lib.iptr(si), len(si), _lib.iptr(sq), len(sq), ctypes.c_double(c)))


def coreDualMisLoop(cp_X, cp_X_rm, cp_Y_X, gam_vb, log_Phi_theta_plus_zeta, log_1_min_Phi_theta_plus_zeta, log_sig2_inv_vb,
                    log_tau_vb, m1_beta, cp_betaX_X, mu_beta_vb, sig2_beta_vb, tau_vb, shuffled_ind, sample_q, c=1.0,
                    device=0):
    """cp_X_rm: list of q Fortran-ordered p x p matrices (R/atlasqtl_global_local_core.R:25-32); sig2_beta_vb p x q.
    In place on gam_vb, m1_beta, cp_betaX_X, mu_beta_vb, like the reference (src/coreLoop.cpp:91-138)."""
    lib = _lib.load()
    p, q = gam_vb.shape
    mats = (cp_X, cp_Y_X, gam_vb, log_Phi_theta_plus_zeta, log_1_min_Phi_theta_plus_zeta, m1_beta, cp_betaX_X, mu_beta_vb,
            sig2_beta_vb) + tuple(cp_X_rm)
    for m in mats:
        if not (isinstance(m, np.ndarray) and m.dtype == np.float64 and m.flags.f_contiguous):
            raise TypeError("matrices must be Fortran-ordered float64 arrays (R's double storage)")
    if cp_X.shape != (p, p) or cp_Y_X.shape != (q, p) or sig2_beta_vb.shape != (p, q):
        raise ValueError("cp_X must be p x p, cp_Y_X q x p and sig2_beta_vb p x q")
    if len(cp_X_rm) != q or any(m.shape != (p, p) for m in cp_X_rm):
        raise ValueError("cp_X_rm must be a list of q p x p matrices")
    vecs = [np.ascontiguousarray(v, dtype=np.float64) for v in (log_tau_vb, tau_vb)]
    si = np.ascontiguousarray(shuffled_ind, dtype=np.int32)
    sq = np.ascontiguousarray(sample_q, dtype=np.int32)
    rm = (_lib._DP * q)(*[_lib.dptr(m) for m in cp_X_rm])
    _lib.check(lib.aq_coreDualMisLoop(ctypes.c_int(device), p, q, _lib.dptr(cp_X), rm, _lib.dptr(cp_Y_X), _lib.dptr(gam_vb),
                                      _lib.dptr(log_Phi_theta_plus_zeta), _lib.dptr(log_1_min_Phi_theta_plus_zeta),
                                      ctypes.c_double(log_sig2_inv_vb), _lib.dptr(vecs[0]), _lib.dptr(m1_beta),
                                      _lib.dptr(cp_betaX_X), _lib.dptr(mu_beta_vb), _lib.dptr(sig2_beta_vb),
                                      _lib.dptr(vecs[1]), _lib.iptr(si), len(si), _lib.iptr(sq), len(sq),
                                      ctypes.c_double(c)))


def logistic_neg_device(x, device=0, variant=0):
    """1 / (1 + exp(x)) evaluated on the device by the sweep's own routine (aq_test_logistic): a test hook.
    variant 0 / 1: the two versions of the routine the kernel configurations use."""
    lib = _lib.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    _lib.check(lib.aq_test_logistic(ctypes.c_int(device), ctypes.c_int(variant), _lib.dptr(x), _lib.dptr(out),
                                    ctypes.c_int(x.size)))
    return out
