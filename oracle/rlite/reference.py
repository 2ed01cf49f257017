"""Run the reference's OWN R sources through the rlite evaluator.  TEST INFRASTRUCTURE ONLY (part of oracle/).

`load()` sources R/*.R from where they lie under /root/reference (never copied into this repository) and binds
`.Call(`_atlasqtl_coreDualLoop`, ...)` / `.Call(`_atlasqtl_coreDualMisLoop`, ...)` -- the two native symbols the
generated R/RcppExports.R glue names -- to the reference's own src/coreLoop.cpp as compiled by oracle/Makefile
(oracle/_ref).  Every R statement executed is therefore the reference's; what stands in for R itself is this
evaluator, and what stands in for nmath / gsl are the SciPy functions named in base.py.

/root/reference does not exist on the GPU box: only tests/golden/make_rlite_golden.py and CPU tests that skip
without it may call this module.
"""
import os

import numpy as np

from .. import native
from .interp import Interp
from .values import Builtin, RError, RList, chr_, from_py, to_py

REF = os.environ.get("ATLASQTL_REFERENCE", "/root/reference")
# > 1: one .Call to coreDualLoop is issued as that many concurrent calls on disjoint ranges of sample_q.  Traits are
# independent inside the loop (column k of every in-place argument is touched by trait k only, src/coreLoop.cpp:58-85),
# so the results are those of the single call; used for the BASELINE-size golden runs only.
THREADS = int(os.environ.get("RLITE_REF_THREADS", "1"))
R_FILES = ("utils.R", "update_vb.R", "elbo.R", "RcppExports.R", "atlasqtl_global_local_core.R",
           "summarise_output.R", "prepare_atlasqtl.R", "set_hyper_init.R", "atlasqtl.R")


def available():
    return os.path.exists(os.path.join(REF, "R", "atlasqtl_global_local_core.R")) and native.ref_available()


def _f64(v, what):
    a = v.a
    if a.dtype != np.float64 or (a.ndim == 2 and not a.flags.f_contiguous):
        raise RError(f".Call: {what} must be a double vector / column-major matrix (in-place update)")
    return a


def _dot_call(it, pos, named):
    sym = pos[0].a[0]
    args = pos[1:]
    if sym == "_atlasqtl_coreDualLoop":
        (cp_X, cp_Y_X, gam, lphi, l1phi, lsig, ltau, m1, cpb, mu, s2b, tau, shuf, sq, c) = args
        fixed = (_f64(cp_X, "cp_X"), _f64(cp_Y_X, "cp_Y_X"), _f64(gam, "gam_vb"), _f64(lphi, "log_Phi"),
                 _f64(l1phi, "log_1_min_Phi"), float(lsig.a[0]), _f64(ltau, "log_tau_vb"), _f64(m1, "m1_beta"),
                 _f64(cpb, "cp_betaX_X"), _f64(mu, "mu_beta_vb"), _f64(s2b, "sig2_beta_vb"), _f64(tau, "tau_vb"),
                 shuf.a.astype(np.int32))
        sample_q = sq.a.astype(np.int32)
        if THREADS > 1 and sample_q.size >= 2 * THREADS:
            from concurrent.futures import ThreadPoolExecutor
            parts = [np.ascontiguousarray(a) for a in np.array_split(sample_q, THREADS)]
            with ThreadPoolExecutor(THREADS) as ex:
                list(ex.map(lambda part: native.core_dual_loop(*fixed, part, c=float(c.a[0]), impl="reference"), parts))
        else:
            native.core_dual_loop(*fixed, sample_q, c=float(c.a[0]), impl="reference")
        return None
    if sym == "_atlasqtl_coreDualMisLoop":
        (cp_X, cp_X_rm, cp_Y_X, gam, lphi, l1phi, lsig, ltau, m1, cpb, mu, s2b, tau, shuf, sq, c) = args
        p, q = gam.a.shape
        stack = np.empty((p, p, q), order="F")
        for k, m in enumerate(cp_X_rm.items):
            stack[:, :, k] = m.a
        native.ref_core_dual_mis_loop(_f64(cp_X, "cp_X"), stack, _f64(cp_Y_X, "cp_Y_X"), _f64(gam, "gam_vb"),
                                      _f64(lphi, "log_Phi"), _f64(l1phi, "log_1_min_Phi"), float(lsig.a[0]),
                                      _f64(ltau, "log_tau_vb"), _f64(m1, "m1_beta"), _f64(cpb, "cp_betaX_X"),
                                      _f64(mu, "mu_beta_vb"), _f64(s2b, "sig2_beta_vb"), _f64(tau, "tau_vb"),
                                      shuf.a.astype(np.int32), sq.a.astype(np.int32), c=float(c.a[0]))
        return None
    raise RError(f".Call: unknown native symbol {sym}")


def load(files=R_FILES, ref=REF):
    it = Interp()
    g = it.globalenv.vars
    g[".Call"] = Builtin(_dot_call, ".Call")
    for sym in ("_atlasqtl_coreDualLoop", "_atlasqtl_coreDualMisLoop"):
        g[sym] = chr_(sym)
    for f in files:
        it.source(os.path.join(ref, "R", f))
    return it


def _copy_in(x):
    if isinstance(x, dict):
        return RList([_copy_in(v) for v in x.values()], list(x.keys()))
    if isinstance(x, np.ndarray):
        return from_py(np.array(x, dtype=np.float64 if x.dtype.kind == "f" else x.dtype, order="F", copy=True))
    return from_py(x)


def with_class(lst, cls):
    lst.attrs = {"class": chr_(cls)}
    return lst


def global_local_core(Y, X, shr_fac_inv, anneal, df, tol, maxit, list_hyper, list_init, it=None, hook=None,
                      **kwargs):
    """atlasqtl_global_local_core_ (R/atlasqtl_global_local_core.R:8-433) on numpy inputs; returns a dict.
    hook(name, value): called with lb_new after every elbo_global_local_ evaluation (wrapped at the R level, the
    reference function itself is untouched)."""
    it = it or load()
    if hook is not None:
        orig = it.globalenv.vars["elbo_global_local_"]

        def traced(it_, pos, named):
            v = it_.apply(orig, pos, named, it_.globalenv)
            hook("lb", float(v.a[0]))
            return v
        it.globalenv.vars["elbo_global_local_"] = Builtin(traced, "elbo_global_local_")
    args = [_copy_in(np.asarray(Y, dtype=np.float64)), _copy_in(np.asarray(X, dtype=np.float64)),
            from_py(float(shr_fac_inv)), None if anneal is None else from_py(np.asarray(anneal, dtype=np.float64)),
            from_py(float(df)), from_py(float(tol)), from_py(float(maxit)), from_py(0.0),
            _copy_in(list_hyper), _copy_in(list_init)]
    named = {k: from_py(v) for k, v in kwargs.items()}
    try:
        out = it.call("atlasqtl_global_local_core_", *args, **named)
    finally:
        if hook is not None:
            it.globalenv.vars["elbo_global_local_"] = orig
    return to_py(out)
