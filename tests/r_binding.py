"""Harness that executes the R-side binding (bindings/R/atlasqtl_b200_core.R) -- TEST INFRASTRUCTURE.

The R file is run by the R evaluator of oracle/rlite.  Two things it needs from its surroundings are supplied here:

* `.Call(`_atlasqtl_aq_*`, ...)`: an emulation of bindings/R/atlasqtl_b200_shim.c (which cannot be compiled without R
  headers): the same symbols, argument order, argument checks (double matrix / vector of the context's dimensions,
  scalars) and returned list names, forwarding to a context object -- the real `atlasqtl_b200.device.SweepContext`
  (ctypes over the C ABI, on a GPU) or the oracle-backed test double (CPU);
* the package's own p-, q- and scalar-sized helpers the binding calls unchanged (get_annealing_ladder_,
  update_*_vb_, Q_approx_vec, e_*_ ...): taken from the reference's R files where /root/reference exists
  (`helpers="reference"`), else NumPy stand-ins from oracle/vb_oracle.py (`helpers="standin"`, the GPU box).
"""
import os

import numpy as np

from oracle import vb_oracle
from oracle.rlite import parser as P
from oracle.rlite.interp import Interp
from oracle.rlite.values import Builtin, RError, RList, V, chr_, dbl, from_py, lgl, to_py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINDING = os.path.join(ROOT, "bindings", "R", "atlasqtl_b200_core.R")
REF_HELPER_FILES = ("utils.R", "update_vb.R", "elbo.R")

SYMBOLS = ("_atlasqtl_aq_create", "_atlasqtl_aq_destroy", "_atlasqtl_aq_set_order", "_atlasqtl_aq_set_state",
           "_atlasqtl_aq_get_state", "_atlasqtl_aq_refresh_tables", "_atlasqtl_aq_sweep", "_atlasqtl_aq_rowsums_zpart",
           "_atlasqtl_aq_set_missing", "_atlasqtl_aq_set_state_mis", "_atlasqtl_aq_sweep_mis")


class ExtPtr:
    """R external pointer to an aq_ctx."""

    def __init__(self, ctx):
        self.ctx = ctx


def _need_mat(v, name, nrow, ncol):
    if not (isinstance(v, V) and v.a.dtype == np.float64 and v.a.ndim == 2 and v.a.shape == (nrow, ncol)):
        raise RError(f"{name} must be a {nrow} x {ncol} double matrix")
    return v.a


def _need_vec(v, name, n):
    if not (isinstance(v, V) and v.a.dtype == np.float64 and v.a.size == n):
        raise RError(f"{name} must be a double vector of length {n}")
    return v.a.reshape(-1)


def _need_scalar(v, name):
    if not (isinstance(v, V) and v.a.dtype.kind in "fi" and v.a.size == 1):
        raise RError(f"{name} must be a numeric scalar")
    return float(v.a.reshape(-1)[0])


def _qlist(d, names):
    return RList([dbl(np.asarray(d[k], dtype=np.float64)) for k in names], list(names))


class ShimEmulation:
    def __init__(self, context_factory):
        self.factory = context_factory
        self.calls = {}
        self.live = 0

    def ctx(self, ptr):
        if not isinstance(ptr, ExtPtr) or ptr.ctx is None:
            raise RError("invalid or destroyed atlasqtl_b200 context")
        return ptr.ctx

    def __call__(self, it, pos, named):
        sym = pos[0].a[0]
        self.calls[sym] = self.calls.get(sym, 0) + 1
        a = pos[1:]
        if sym == "_atlasqtl_aq_create":
            X, Y, device = a
            if not (isinstance(X, V) and isinstance(Y, V) and X.a.dtype == np.float64 and Y.a.dtype == np.float64
                    and X.a.ndim == 2 and Y.a.ndim == 2):
                raise RError("X and Y must be double matrices")
            if X.a.shape[0] != Y.a.shape[0]:
                raise RError("X and Y must have the same number of rows")
            if device.a.dtype.kind != "i":
                raise RError("device must be an integer")
            self.live += 1
            return ExtPtr(self.factory(X.a, Y.a))
        c = self.ctx(a[0])
        n, p, q = c.n, c.p, c.q
        if sym == "_atlasqtl_aq_destroy":
            c.close()
            a[0].ctx = None
            self.live -= 1
            return None
        if sym == "_atlasqtl_aq_set_order":
            c.set_order(None if a[1] is None else a[1].a.astype(np.int32))
            return None
        if sym == "_atlasqtl_aq_set_state":
            return _qlist(c.set_state(_need_mat(a[1], "gam_vb", p, q), _need_mat(a[2], "mu_beta_vb", p, q)),
                          ("colsum_gam", "colsum_gam_mu2", "colsum_beta2", "resid_sq"))
        if sym == "_atlasqtl_aq_get_state":
            outs = [None if v is None else _need_mat(v, nm, p, q) for v, nm in zip(a[1:4], ("gam_vb", "mu_beta_vb", "beta_vb"))]
            st = c.get_state(gam=outs[0] is not None, mu=outs[1] is not None, beta=outs[2] is not None)
            for dst, key in zip(outs, ("gam_vb", "mu_beta_vb", "beta_vb")):
                if dst is not None:
                    dst[...] = st[key]     # in place on the caller's matrix, like the C shim
            return None
        if sym == "_atlasqtl_aq_refresh_tables":
            want = isinstance(a[4], V) and a[4].a.dtype.kind == "b" and bool(a[4].a.reshape(-1)[0])
            part = c.refresh_tables(_need_vec(a[1], "theta_vb", p), _need_vec(a[2], "zeta_vb", q),
                                    c_next=_need_scalar(a[3], "c_next"), want_elbo=want)
            return dbl(np.nan if part is None else part)
        if sym == "_atlasqtl_aq_sweep":
            s = c.sweep(_need_scalar(a[1], "c"), _need_scalar(a[2], "log_sig2_inv_vb"), _need_vec(a[3], "tau_vb", q),
                        _need_vec(a[4], "log_tau_vb", q), _need_vec(a[5], "sig2_beta_vb", q))
            return _qlist(s, ("colsum_gam", "colsum_gam_mu2", "colsum_beta2", "resid_sq", "colsum_zpart"))
        if sym == "_atlasqtl_aq_rowsums_zpart":
            if a[1] is not None and int(_need_scalar(a[1], "p")) != p:
                raise RError(f"p does not match the context ({p})")
            return dbl(c.rowsums_zpart())
        if sym == "_atlasqtl_aq_set_missing":
            return dbl(c.set_missing(_need_mat(a[1], "mis_pat", n, q)))
        if sym == "_atlasqtl_aq_set_state_mis":
            return _qlist(c.set_state_mis(_need_mat(a[1], "gam_vb", p, q), _need_mat(a[2], "mu_beta_vb", p, q)),
                          ("colsum_gam", "colsum_gam_mu2", "colsum_beta2", "resid_sq", "colsum_xn_gam",
                           "colsum_xn_gam_mu2", "colsum_xn_beta2"))
        if sym == "_atlasqtl_aq_sweep_mis":
            s = c.sweep_mis(_need_scalar(a[1], "c"), _need_scalar(a[2], "log_sig2_inv_vb"),
                            _need_scalar(a[3], "sig2_inv_vb"), _need_vec(a[4], "tau_vb", q), _need_vec(a[5], "log_tau_vb", q))
            return _qlist(s, ("colsum_gam", "colsum_gam_mu2", "colsum_sig2b_gam", "colsum_xn_gam_mu2",
                              "colsum_xn_sig2b_gam", "colsum_xn_beta2", "resid_sq", "colsum_zpart", "colsum_gam_logsig2b"))
        raise RError(f".Call: unknown symbol {sym}")


# ------------------------------------------------------------------------------------------------ stand-in helpers
def _num(v):
    a = np.asarray(v.a, dtype=np.float64)
    return float(a.reshape(-1)[0]) if a.size == 1 else a.reshape(-1)


def _wrap(fn, defaults=()):
    """NumPy function -> R builtin: positional / named numeric arguments, defaults as (name, value) after the positionals."""
    def f(it, pos, named):
        args = [None if v is None else _num(v) for v in pos]
        kw = {k: _num(v) for k, v in named.items()}
        for name, val in defaults:
            kw.setdefault(name, val)
        return dbl(np.asarray(fn(*args, **kw), dtype=np.float64))
    return f


def _named_list_special(it, env, args):
    """create_named_list_(a, b, ...): list(a = a, b = b, ...) -- names are the argument expressions."""
    return RList([it.eval(ex, env) for _, ex in args], [P.deparse(ex) for _, ex in args])


def install_standin_helpers(it):
    g = it.globalenv.vars
    o = vb_oracle
    g["get_annealing_ladder_"] = Builtin(lambda it_, pos, named: dbl(o.get_annealing_ladder_(tuple(_num(pos[0])))),
                                         "get_annealing_ladder_")
    g["update_sig2_c0_vb_"] = Builtin(_wrap(lambda d, s02, c=1.0: o.update_sig2_c0_vb_(d, s02, c)), "update_sig2_c0_vb_")
    g["update_nu_vb_"] = Builtin(_wrap(lambda nu, sum_gam, c=1.0: o.update_nu_vb_(nu, sum_gam, c)), "update_nu_vb_")
    g["update_log_tau_vb_"] = Builtin(_wrap(o.update_log_tau_vb_), "update_log_tau_vb_")
    g["update_log_sig2_inv_vb_"] = Builtin(_wrap(o.update_log_sig2_inv_vb_), "update_log_sig2_inv_vb_")
    g["update_annealed_lam2_inv_vb_"] = Builtin(_wrap(lambda L, c, df: o.update_annealed_lam2_inv_vb_(L, c, int(df))),
                                                "update_annealed_lam2_inv_vb_")
    g["Q_approx_vec"] = Builtin(_wrap(o.Q_approx_vec), "Q_approx_vec")
    g["e_tau_"] = Builtin(_wrap(o.e_tau_), "e_tau_")
    g["e_theta_hs_"] = Builtin(_wrap(lambda *a: o.e_theta_hs_(*a[:-1], int(a[-1]))), "e_theta_hs_")
    g["e_zeta_"] = Builtin(_wrap(o.e_zeta_), "e_zeta_")
    g["e_sig2_inv_"] = Builtin(_wrap(o.e_sig2_inv_), "e_sig2_inv_")
    g["e_sig2_inv_hs_"] = Builtin(_wrap(o.e_sig2_inv_hs_), "e_sig2_inv_hs_")
    g["create_named_list_"] = Builtin(_named_list_special, "create_named_list_", special=True)
    g["checkpoint_clean_up_"] = Builtin(lambda it_, pos, named: None, "checkpoint_clean_up_")

    def no_checkpoint(it_, pos, named):
        raise RError("checkpoint_ needs the package's own R/utils.R")
    g["checkpoint_"] = Builtin(no_checkpoint, "checkpoint_")


def load(context_factory=None, helpers="standin", shim="emulation"):
    """-> (interpreter with atlasqtl_b200_core_ defined, the `.Call` object).  shim = "emulation": the Python emulation
    of the C shim over context_factory; shim = "real": the shipped C shim itself, compiled against the stub R runtime
    (tests/r_shim_real.py) and linked to libatlasqtl_b200.so (needs a GPU to get past aq_create)."""
    it = Interp()
    if shim == "real":
        from r_shim_real import RealShim
        shim = RealShim()
    else:
        shim = ShimEmulation(context_factory)
    g = it.globalenv.vars
    g[".Call"] = Builtin(shim, ".Call")
    for sym in SYMBOLS:
        g[sym] = chr_(sym)
    if helpers == "reference":
        from oracle.rlite import reference as R
        for f in REF_HELPER_FILES:
            it.source(os.path.join(R.REF, "R", f))
    elif helpers == "standin":
        install_standin_helpers(it)
    else:
        raise ValueError(helpers)
    it.source(BINDING)
    return it, shim


def run_core(it, Y, X, anneal, tol, hyper, init, maxit=1000, thinned=True, full_output=False, trace=None):
    """atlasqtl_b200_core_ on NumPy inputs -> dict.  trace: list receiving (it, ELBO) of every evaluation (lb_hook)."""
    q = Y.shape[1]

    def hook(it_, pos, named):
        if trace is not None:
            trace.append((int(pos[0].a[0]), float(pos[1].a[0])))
        return None
    out = it.call("atlasqtl_b200_core_", from_py(np.array(Y, dtype=np.float64, order="F")),
                  from_py(np.array(X, dtype=np.float64, order="F")), from_py(float(q)),
                  None if anneal is None else from_py(np.asarray(anneal, dtype=np.float64)), from_py(1.0),
                  from_py(float(tol)), from_py(float(maxit)), from_py(0.0),
                  RList([from_py(np.asarray(v, dtype=np.float64)) for v in hyper.values()], list(hyper.keys())),
                  RList([from_py(np.array(v, dtype=np.float64, order="F")) for v in init.values()], list(init.keys())),
                  thinned_elbo_eval=lgl(bool(thinned)), debug=lgl(True), full_output=lgl(bool(full_output)),
                  lb_hook=Builtin(hook, "lb_hook"))
    return to_py(out)
