"""`coreDualLoop`: the reference's R closure (R/RcppExports.R:4-6) with the same 15 arguments and the same
in-place outputs, backed by the CUDA sweep (aq_coreDualLoop, include/atlasqtl_b200.h)."""
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
