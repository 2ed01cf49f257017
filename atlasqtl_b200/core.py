"""VB outer loop of the global-local (horseshoe) model: host-side mirror of
`atlasqtl_global_local_core_` (reference R/atlasqtl_global_local_core.R:8-433) driving the CUDA
sweep through the C ABI.

What moved to the device (include/atlasqtl_b200.h) and what stays here:

  device  step 10 (coreDualLoop), m2_beta / Z / log-CDF tables (steps 11, 12, 19), every p x q and
          n x q reduction (steps 1-5, 16, 18) and the p x q part of ELBO term B;
  host    the p-, q- and scalar-sized algebra between those calls -- the same statements as
          R/atlasqtl_global_local_core.R:134-150, :241-290, :318-375 and R/elbo.R, written against
          the per-trait / per-SNP sums the device hands back.

Traits are independent inside a sweep, so several processes can each own a slab of traits
(`slab`): the only cross-slab quantities are rowSums(Z) (p), a few scalars per sweep and the ELBO
partial sums; they go through `comm.allreduce_sum` (NCCL via torch.distributed, see dist.py).
"""
import math
import time
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
from scipy import special as sp

from .device import PreparedPredictors, SweepContext

ALL_EQUAL_TOL = 1.5e-8  # isTRUE(all.equal(c, 1))


class SerialComm:
    """Single-slab stand-in for dist.TorchComm."""
    rank, world_size = 0, 1

    def allreduce_sum(self, x):
        return x

    def allreduce_min(self, x):
        return x


# ----------------------------------------------------------------------------- small host helpers
_POOL = None
_BG = None


def _background():
    """One host thread for work that overlaps the (GIL-releasing) sweep call."""
    global _BG
    if _BG is None:
        _BG = ThreadPoolExecutor(1)
    return _BG


def _host_threads():
    """Host threads this process may use for the p-vector special functions: its share of the box's cores when several
    ranks run side by side (torchrun exports LOCAL_WORLD_SIZE); 8 ranks x 32 threads on a 32-core host made every rank
    wait for the horseshoe update (round 1: hs_wait 1.1 ms per iteration at 8 GPUs)."""
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    share = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    return max(1, min(32, ncpu // share))


def _pmap(fn, x, min_chunk=1024):
    """Apply an element-wise SciPy special function over a long p-vector on all host threads (the ufunc inner loops
    release the GIL).  These are the only non-trivial host costs per iteration: exp1 / gammaincc at p = 50k take
    tens of milliseconds single-threaded, comparable to the GPU sweep once the traits are spread over 8 GPUs."""
    global _POOL
    x = np.ascontiguousarray(x)
    nthr = min(_host_threads(), max(1, x.size // min_chunk))
    if nthr <= 1:
        return fn(x)
    if _POOL is None:
        _POOL = ThreadPoolExecutor(_host_threads())
    out = np.empty_like(x, dtype=np.float64)
    bounds = np.linspace(0, x.size, nthr + 1).astype(int)

    def work(i):
        out[bounds[i]:bounds[i + 1]] = fn(x[bounds[i]:bounds[i + 1]])

    list(_POOL.map(work, range(nthr)))
    return out


def get_annealing_ladder_(anneal):
    """R/utils.R:108-146."""
    k_m = 1.0 / anneal[1]
    m = int(anneal[2])
    down = np.arange(m, 0, -1, dtype=np.float64)
    if anneal[0] == 1:
        return (1 + (k_m ** (1.0 / (1 - m)) - 1)) ** (1 - down)
    if anneal[0] == 2:
        return 1 / (1 + ((1 / k_m - 1) / (m - 1)) * (down - 1))
    return k_m + ((1 - k_m) / (m - 1)) * (np.arange(1, m + 1, dtype=np.float64) - 1)


def check_annealing_(anneal):
    """R/prepare_atlasqtl.R:100-124."""
    if anneal is None:
        return
    if len(anneal) != 3:
        raise ValueError("anneal must be NULL/None or a vector of length 3.")
    if anneal[0] not in (1, 2, 3):
        raise ValueError("The annealing spacing scheme must be set to 1 for geometric 2 for harmonic or 3 for "
                         "linear spacing.")
    if anneal[1] < 1.5:
        raise ValueError("Initial annealing temperature very small. May not be large enough for a successful "
                         "exploration. Please increase it or select no annealing.")
    if anneal[2] > 1000 or anneal[2] != int(anneal[2]) or anneal[2] < 1:
        raise ValueError("Temperature grid size must be a natural number <= 1000.")


def _lentz_vecwide(xu, eps1, eps2):
    """Modified Lentz continued fraction of R/utils.R:393-418.  The reference stops on the VECTOR-WIDE criterion
    max|Delta - 1| < eps2, so every element's value depends on the whole vector; kept as one vectorised loop
    (~25 iterations; threading it costs more in synchronisation than it saves)."""
    f_p = np.full_like(xu, eps1)
    C_p = np.full_like(xu, eps1)
    D_p = np.zeros_like(xu)
    Delta = np.full_like(xu, 2 + eps2)
    j = 1
    while np.max(np.abs(Delta - 1)) >= eps2:
        j += 1
        D_c = 1 / (xu + 2 * j - 1 - ((j - 1) ** 2) * D_p)
        C_c = xu + 2 * j - 1 - ((j - 1) ** 2) / C_p
        Delta = C_c * D_c
        f_p = f_p * Delta
        C_p, D_p = C_c, D_c
    return f_p


def Q_approx_vec(x, eps1=1e-30, eps2=1e-7, only=None):
    """E1(x) exp(x): gsl::expint_E1 branch for x <= 1, modified Lentz for x > 1 with the reference's
    vector-wide stopping rule (R/utils.R:380-423).  only = (j0, j1): the element-wise E1 branch is evaluated for the
    entries [j0, j1) alone (the others are left at 0); the Lentz branch always sees the whole vector, as its stopping
    rule demands."""
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    lo = x <= 1
    if only is not None:
        inside = np.zeros(x.shape, dtype=bool)
        inside[only[0]:only[1]] = True
        lo_eval = lo & inside
    else:
        lo_eval = lo
    if lo_eval.any():
        out[lo_eval] = _pmap(lambda v: sp.exp1(v) * np.exp(v), x[lo_eval])
    if (~lo).any():
        xu = np.ascontiguousarray(x[~lo])
        out[~lo] = 1 / (xu + 1 + _lentz_vecwide(xu, eps1, eps2))
    return out


def _upper_gamma(a, x):
    return sp.gamma(a) * _pmap(lambda v: sp.gammaincc(a, v), x)  # gsl::gamma_inc(a, x), a > 0 on this path


def update_annealed_lam2_inv_vb_(L_vb, c, df):
    if df != 1:
        raise NotImplementedError("df is fixed to 1 by the reference (R/atlasqtl.R:272)")
    return _upper_gamma(2 - c, L_vb) / (_upper_gamma(1 - c, L_vb) * L_vb) - 1  # R/update_vb.R:74


def _e_sig2_inv(nu, nu_vb, log_sig2_inv_vb, rho, rho_vb, sig2_inv_vb):
    return ((nu - nu_vb) * log_sig2_inv_vb - (rho - rho_vb) * sig2_inv_vb + nu * math.log(rho)
            - nu_vb * math.log(rho_vb) - sp.gammaln(nu) + sp.gammaln(nu_vb))  # R/elbo.R:41-46


# ----------------------------------------------------------------------------- checkpoints (R/utils.R:571-627)
def _checkpoint_file(checkpoint_path, it, rank, world_size):
    tag = "" if world_size == 1 else f"_slab{rank}"
    return f"{checkpoint_path}tmp_output_it_{it}{tag}.npz"


def checkpoint_(it, checkpoint_path, ctx, theta_vb, zeta_vb, converged, lb_new, lb_old, lam2_inv_vb=None,
                sig02_inv_vb=None, rate=100, comm=None, extra=None):
    """Every `rate` iterations: the device -> host hand-off of gam_vb / beta_vb (aq_get_state) and the fields the
    reference saves (`tmp_vb`, R/utils.R:596-602), keeping the last two files (:604-607).  One file per trait slab.
    `extra` adds what a RESUME needs (mu_beta_vb, tau_vb, ...), which the reference's checkpoints lack."""
    if checkpoint_path is None or it % rate != 0:
        return None
    rank, world = (comm.rank, comm.world_size) if comm is not None else (0, 1)
    fields = dict(theta_vb=np.array(theta_vb), zeta_vb=np.array(zeta_vb), converged=converged, it=it, lb_new=lb_new,
                  diff_lb=abs(lb_new - lb_old))
    if lam2_inv_vb is not None:
        fields["lam2_inv_vb"] = np.array(lam2_inv_vb)
    if sig02_inv_vb is not None:
        fields["sig02_inv_vb"] = sig02_inv_vb
    if extra is not None:
        fields.update({k: np.array(v) for k, v in extra.items()})
    path = _checkpoint_file(checkpoint_path, it, rank, world)
    old = _checkpoint_file(checkpoint_path, it - 2 * rate, rank, world)

    def write(fetch):
        st = fetch(gam=True, mu=extra is not None, beta=True)
        fields.update(beta_vb=st["beta_vb"], gam_vb=st["gam_vb"])
        if extra is not None:
            fields["mu_beta_vb"] = st["mu_beta_vb"]
        np.savez(path, **fields)
        if os.path.exists(old):
            os.remove(old)
        return path

    if hasattr(ctx, "snapshot"):
        # device path: freeze the state in stream order (milliseconds), then transpose / download / write on a host thread
        # and a copy stream while the sweeps go on (include/atlasqtl_b200.h, aq_snapshot / aq_snapshot_fetch)
        checkpoint_join_(ctx)
        ctx.snapshot()
        ctx._ckpt_future = _checkpoint_pool().submit(write, ctx.snapshot_fetch)
        return path
    return write(ctx.get_state)


_CKPT = None


def _checkpoint_pool():
    global _CKPT
    if _CKPT is None:
        _CKPT = ThreadPoolExecutor(1)
    return _CKPT


def checkpoint_join_(ctx):
    """Wait for a checkpoint still being written in the background (re-raises what it raised)."""
    fut = getattr(ctx, "_ckpt_future", None)
    if fut is not None:
        ctx._ckpt_future = None
        fut.result()


def checkpoint_clean_up_(checkpoint_path, comm=None):
    """R/utils.R:612-625: remove the temporary files once the run has finished."""
    if checkpoint_path is None:
        return
    import glob
    rank, world = (comm.rank, comm.world_size) if comm is not None else (0, 1)
    tag = "" if world == 1 else f"_slab{rank}"
    for f in glob.glob(f"{checkpoint_path}tmp_output_it_*{tag}.npz"):
        os.remove(f)


def init_from_checkpoint(path, list_init):
    """Extension (the reference cannot resume): a list_init that restarts the run from a checkpoint written with the
    resume fields.  theta_vb, zeta_vb, gam_vb, mu_beta_vb, tau_vb, sig2_beta_vb, sig2_theta_vb, sig02_inv_vb are taken
    from the file, everything else from `list_init`.

    Checkpoints are written after annealing only (like the reference's), so the restart must run with `anneal=None`:
    the iteration counter starts again at 1 and the ELBO-thinning schedule with it, i.e. the restart continues the
    same fixed-point iteration but is not a bit-for-bit replay of the original run's bookkeeping.  Runs with missing
    responses are refused: their checkpoint holds the q-vector part of sig2_beta_vb only, while the state the first
    restarted iteration needs (m2_beta with the p x q sig2_beta_vb of the saved sweep) cannot be rebuilt from it."""
    d = np.load(path)
    if "has_missing" in d.files and bool(d["has_missing"]):
        raise ValueError("resuming from a checkpoint of a run with missing responses is not supported")
    out = dict(list_init)
    for key in ("gam_vb", "mu_beta_vb", "theta_vb", "zeta_vb", "tau_vb", "sig2_beta_vb", "sig2_theta_vb"):
        out[key] = np.array(d[key])
    out["sig02_inv_vb"] = float(d["sig02_inv_vb"])
    return out


# ----------------------------------------------------------------------------- the core
def atlasqtl_global_local_core_(Y, X, shr_fac_inv, anneal, df, tol, maxit, verbose, list_hyper, list_init,
                                checkpoint_path=None, trace_path=None, full_output=False,
                                thinned_elbo_eval=True, debug=False, batch="y", *, comm=None, slab=None,
                                device=0, context_factory=None, order_fn=None, trace=None, ctx=None, iter_hook=None,
                                checkpoint_rate=100, keep_checkpoints=False, release_x=None):
    """Same positional arguments as the reference core (R/atlasqtl_global_local_core.R:8-13).

    Y is THIS process's slab of responses (all of Y when comm is None); `slab` = (k_first, k_last)
    gives its position among the q_total = shr_fac_inv traits so that per-trait hyper / init vectors
    (length q_total) can be sliced.  `order_fn(it, p)` supplies shuffled_ind per iteration (identity
    by default, like the reference, :162).  Keyword-only arguments are extensions; the R-facing ones
    keep their meaning.  Missing values in Y (NaN) select the coreDualMisLoop path (:19-38, :172-175), n <= 2048.
    release_x: free the context's untiled copy of X once the state is set up (default: only for a context this call owns).
    """
    if batch != "y":
        raise ValueError("Batch scheme not defined. Exit.")  # only the C++ path is replaced (:179-232)
    mis_pat = None
    Y_raw = Y  # a PreparedPredictors X centres the raw responses on the device (R/prepare_atlasqtl.R:83)
    if np.isnan(Y).any():  # :19-33 (X_norm_sq and the per-trait Gram corrections live on the device)
        mis_pat = np.where(np.isnan(Y), 0.0, 1.0)
        Y = np.where(np.isnan(Y), 0.0, Y)
    if trace_path is not None:
        raise NotImplementedError("trace_path (diagnostic plots of the hotspot variances) is outside this path")
    comm = comm or SerialComm()
    n, q = Y.shape
    p = X.shape[1]
    q_total = int(shr_fac_inv)
    k_first, k_last = slab if slab is not None else (0, q)
    if k_last - k_first != q:
        raise ValueError("slab does not match the number of columns of Y")
    sl = slice(k_first, k_last)

    def per_trait(v):
        v = np.asarray(v, dtype=np.float64)
        return v[sl].copy() if v.shape == (q_total,) else np.full(q, float(v))

    h = list_hyper
    eta, kappa, n0 = per_trait(h["eta"]), per_trait(h["kappa"]), per_trait(h["n0"])
    nu, rho, t02 = float(h["nu"]), float(h["rho"]), float(h["t02"])
    m0, A2_inv = float(h.get("m0", 0.0)), float(h.get("A2_inv", 1.0))

    gam0 = np.asarray(list_init["gam_vb"])
    mu0 = np.asarray(list_init["mu_beta_vb"])
    if gam0.shape[1] == q_total and q_total != q:
        gam0, mu0 = gam0[:, sl], mu0[:, sl]
    sig02_inv_vb = float(list_init["sig02_inv_vb"])
    sig2_beta_vb = per_trait(list_init["sig2_beta_vb"])
    sig2_theta_vb = np.array(list_init["sig2_theta_vb"], dtype=np.float64)
    tau_vb = per_trait(list_init["tau_vb"])
    theta_vb = np.array(list_init["theta_vb"], dtype=np.float64)
    zeta_vb = per_trait(list_init["zeta_vb"])

    anneal_scale = True  # :71
    if anneal is None:
        annealing, c, c_s, it_init, ladder = False, 1.0, 1.0, 1, None
    else:
        annealing = True
        ladder = get_annealing_ladder_(anneal)
        c = float(ladder[0])
        c_s = c if anneal_scale else 1.0
        it_init = int(anneal[2])
    eps = np.finfo(np.float64).eps ** 0.5  # :85
    if thinned_elbo_eval:
        times_conv_sched, batch_conv_sched = np.array([1.0, 5.0, 10.0, 50.0]), [1, 10, 25, 50]
    else:
        times_conv_sched, batch_conv_sched = np.array([1.0]), [1]
    ind_batch_conv = len(batch_conv_sched) + 1
    batch_conv = 1

    t02_inv = 1 / t02
    sig2_zeta_vb = 1 / (c * (p + t02_inv))  # update_sig2_c0_vb_(p, t02, c), :105
    vec_sum_log_det_zeta = -q_total * (math.log(t02) + math.log(p + t02_inv))  # :107
    nu_xi_inv_vb = 1.0  # :119

    own_ctx = ctx is None
    if own_ctx:
        if isinstance(X, PreparedPredictors):
            ctx = X.context(Y_raw)
        else:
            factory = context_factory or (lambda X_, Y_: SweepContext(X_, Y_, device=device))
            ctx = factory(X, Y)
    del Y_raw
    try:
        order = None
        if order_fn is not None:
            order = np.ascontiguousarray(order_fn(1, p), dtype=np.int32)
        ctx.set_order(order)
        if mis_pat is None:
            n_eff = n
            sums = ctx.set_state(gam0, mu0)  # beta_vb, residual, and the sums m2_beta / kappa_vb need (:112-115)
        else:
            n_eff = ctx.set_missing(mis_pat)  # colSums(mis_pat): update_eta_vb_ R/update_vb.R:131, e_y_ R/elbo.R:141
            sums = ctx.set_state_mis(gam0, mu0)
            # m2_beta <- update_m2_beta_(..., sweep = TRUE) with the q-vector sig2_beta_vb of the init (:113)
            sums["colsum_m2"] = sums["colsum_gam_mu2"] + sig2_beta_vb * sums["colsum_gam"]
            sums["colsum_xn_m2"] = sums["colsum_xn_gam_mu2"] + sig2_beta_vb * sums["colsum_xn_gam"]
        del gam0, mu0
        if release_x is None:
            release_x = own_ctx   # a caller-owned context may be reused with another order / with missing responses
        if release_x and order_fn is None and mis_pat is None and hasattr(ctx, "release_x"):
            ctx.release_x()  # identity order for the whole run (:162) and no NA kernel: the tiled X is all the sweeps read
        ctx.refresh_tables(theta_vb, zeta_vb, c_next=c)  # :61-63
        sig2_beta_for_m2 = sig2_beta_vb  # m2_beta always pairs gam/mu with the sig2_beta_vb of their sweep (:113,:235)

        def m2_of(sm):  # colSums(m2_beta)
            return sm["colsum_m2"] if "colsum_m2" in sm else sm["colsum_gam_mu2"] + sig2_beta_for_m2 * sm["colsum_gam"]

        def kappa_rate(sm, s_inv):  # the bracket of update_kappa_vb_ (R/update_vb.R:144-146 / :149-154)
            if mis_pat is None:
                return sm["resid_sq"] + (n - 1 + s_inv) * m2_of(sm) - (n - 1) * sm["colsum_beta2"]
            return sm["resid_sq"] + s_inv * m2_of(sm) + sm["colsum_xn_m2"] - sm["colsum_xn_beta2"]

        glob = comm.allreduce_sum(np.array([sums["colsum_gam"].sum(), np.dot(tau_vb, m2_of(sums)), zeta_vb.sum()]))
        sum_gam, tau_dot_m2, sum_zeta = (float(v) for v in glob)

        converged = False
        lb_new = lb_old = -np.inf
        it = 0
        Q_app = None
        lam2_inv_vb = None

        while (not converged) and (it < maxit):
            lb_old = lb_new
            it += 1
            if iter_hook is not None:
                iter_hook(it, ctx)
            _t = [time.perf_counter()]
            seg = {}

            def _lap(name):  # host wall-clock per segment of the iteration (rec["host_ms"], for profiling the loop itself)
                t1 = time.perf_counter()
                seg[name] = seg.get(name, 0.0) + 1e3 * (t1 - _t[0])
                _t[0] = t1
            if verbose != 0 and comm.rank == 0 and (it == 1 or it % max(5, batch_conv) == 0):
                print(f"Iteration {it}... ")

            nu_vb = c * (nu + sum_gam / 2) - c + 1  # :134
            rho_vb = c * (rho + tau_dot_m2 / 2)  # :135 (uses the tau_vb of the previous iteration)
            sig2_inv_vb = nu_vb / rho_vb  # :137
            eta_vb = c * (eta + n_eff / 2 + sums["colsum_gam"] / 2) - c + 1  # :141
            kappa_vb = c * (kappa + kappa_rate(sums, sig2_inv_vb) / 2)  # :142
            tau_vb = eta_vb / kappa_vb  # :145
            sig2_beta_vb = 1 / (c * (n - 1 + sig2_inv_vb) * tau_vb)  # :147 (p x q, formed on the device, if Y has NAs)
            log_tau_vb = sp.digamma(eta_vb) - np.log(kappa_vb)  # :149
            log_sig2_inv_vb = float(sp.digamma(nu_vb) - math.log(rho_vb))  # :150

            if order_fn is not None and it > 1:
                ctx.set_order(np.ascontiguousarray(order_fn(it, p), dtype=np.int32))  # :160-163

            # The horseshoe scale update (:241-254) reads only the PREVIOUS theta_vb / sig2_theta_vb / sig02_inv_vb, so its
            # p-vector special functions (incomplete gammas, E1 / Lentz) run on a host thread while the GPU sweeps; the
            # statements and their operands are unchanged, only the wall-clock position moves.  With several ranks every
            # rank evaluates the element-wise special functions on ITS slice of the p SNPs only and the slices are
            # exchanged in the sweep's all-reduce (a sum with zeros elsewhere: exact, identical on all ranks) -- 8 ranks
            # each computing all p values on a shared host made every rank wait for this update once the sweep took 18 ms.
            # The Lentz branch is evaluated on the whole vector by everybody: its stopping rule is vector-wide (R/utils.R:402).
            j0, j1 = (comm.rank * p) // comm.world_size, ((comm.rank + 1) * p) // comm.world_size

            def _hs_scale(c_s=c_s, theta=theta_vb, s2t=sig2_theta_vb, s02=sig02_inv_vb, ann=annealing and anneal_scale):
                th2_ = theta ** 2 + s2t - 2 * theta * m0 + m0 ** 2
                L_ = c_s * s02 * shr_fac_inv * th2_ / 2 / df  # :241
                part = np.zeros(p)
                if ann:
                    part[j0:j1] = update_annealed_lam2_inv_vb_(L_[j0:j1], c_s, df)  # :246
                else:
                    part[j0:j1] = Q_approx_vec(L_, only=(j0, j1))[j0:j1]  # :250
                return L_, ann, part
            hs_future = _background().submit(_hs_scale)

            _lap("pre")
            # ---- the sweep (:167-170) with the fused reductions
            if mis_pat is None:
                sums = ctx.sweep(c, log_sig2_inv_vb, tau_vb, log_tau_vb, sig2_beta_vb)
            else:  # coreDualMisLoop (:172-175)
                sums = ctx.sweep_mis(c, log_sig2_inv_vb, sig2_inv_vb, tau_vb, log_tau_vb)
                sums["colsum_m2"] = sums["colsum_gam_mu2"] + sums["colsum_sig2b_gam"]  # update_m2_beta_ with p x q sig2_beta_vb
                sums["colsum_xn_m2"] = sums["colsum_xn_gam_mu2"] + sums["colsum_xn_sig2b_gam"]
            sig2_beta_for_m2 = sig2_beta_vb
            _lap("sweep")
            colsum_m2 = m2_of(sums)  # :235
            local = [sums["colsum_gam"].sum(), np.dot(tau_vb, colsum_m2)]
            if comm.world_size > 1:  # this rank's slice of the horseshoe p-vector rides in the same message
                L_vb, hs_ann, hs_part = hs_future.result()
                local = np.concatenate([local, hs_part])
            if (comm.world_size > 1 and callable(getattr(comm, "allreduce_sum_device", None))
                    and hasattr(ctx, "rowsums_zpart_dev")):
                # the row sums never visit the host before the all-reduce: NCCL reduces the library's device buffer in
                # place (aq_rowsums_zpart_dev), the two scalars ride in a second, tiny message
                ptr = ctx.rowsums_zpart_dev()
                _lap("rowsums")
                rowsum_zpart = comm.allreduce_sum_device(ptr, p)
                glob = comm.allreduce_sum(np.asarray(local))
                sum_gam, tau_dot_m2, hs_vec = float(glob[0]), float(glob[1]), glob[2:]
            else:
                rows = ctx.rowsums_zpart()
                _lap("rowsums")
                glob = comm.allreduce_sum(np.concatenate([rows, local]))
                rowsum_zpart, sum_gam, tau_dot_m2, hs_vec = glob[:p], float(glob[p]), float(glob[p + 1]), glob[p + 2:]
            _lap("allreduce")

            sqrt_c = 1.0 if abs(c - 1) < ALL_EQUAL_TOL else math.sqrt(c)  # R/update_vb.R:219-229
            rowsums_Z = rowsum_zpart / sqrt_c + q_total * theta_vb + sum_zeta  # :237
            colsums_Z = sums["colsum_zpart"] / sqrt_c + theta_vb.sum() + p * zeta_vb

            rho_xi_inv_vb = c_s * (A2_inv + sig02_inv_vb)  # :242
            if comm.world_size == 1:
                L_vb, hs_ann, hs_vec = hs_future.result()
            _lap("hs_wait")
            if hs_ann:
                lam2_inv_vb = hs_vec  # :246
            else:
                Q_app = hs_vec  # :250
                lam2_inv_vb = 1 / (Q_app * L_vb) - 1  # :254
            xi_inv_vb = nu_xi_inv_vb / rho_xi_inv_vb  # :276
            prior_prec = sig02_inv_vb * lam2_inv_vb * shr_fac_inv
            sig2_theta_vb = 1 / (c * (q_total + prior_prec))  # :278
            theta_vb = c * sig2_theta_vb * (rowsums_Z + prior_prec * m0 - sum_zeta)  # :280
            nu_s0_vb = c_s * (0.5 + p / 2) - c_s + 1  # :283
            rho_s0_vb = c_s * (xi_inv_vb + np.sum(lam2_inv_vb * shr_fac_inv * (
                theta_vb ** 2 + sig2_theta_vb - 2 * theta_vb * m0 + m0 ** 2)) / 2)  # :285
            sig02_inv_vb = float(nu_s0_vb / rho_s0_vb)  # :288
            zeta_vb = c * sig2_zeta_vb * (colsums_Z + t02_inv * n0 - theta_vb.sum())  # :290

            rec = dict(it=it, c=c, annealing=annealing, lb=None, sig2_inv_vb=sig2_inv_vb, sig02_inv_vb=sig02_inv_vb,
                       sum_gam=sum_gam, sweep_ms=getattr(ctx, "last_sweep_ms", lambda: float("nan"))())
            if hasattr(ctx, "last_ms"):
                rec["rows_ms"] = ctx.last_ms(1)
            c_prev = c
            want_elbo = False
            if annealing:  # :318-336
                sig2_zeta_vb = c * sig2_zeta_vb
                c = float(ladder[it]) if it < len(ladder) else 1.0
                c_s = c if anneal_scale else 1.0
                sig2_zeta_vb = sig2_zeta_vb / c
                if abs(c - 1) < ALL_EQUAL_TOL:
                    annealing = False
                    if verbose != 0 and comm.rank == 0:
                        print("** Exiting annealing mode. **\n")
            else:
                want_elbo = it <= it_init + 1 or it % batch_conv == 0 or it % batch_conv == 1  # :342

            _lap("theta_zeta")
            # theta / zeta changed: refresh D, W, I0 for the next sweep (:293-295), ELBO-B part on demand
            elbo_b_dev = ctx.refresh_tables(theta_vb, zeta_vb, c_next=c, want_elbo=want_elbo)
            _lap("tables")
            sum_zeta_local = zeta_vb.sum()
            if hasattr(ctx, "last_ms"):
                rec["tables_ms"] = ctx.last_ms(2)

            if want_elbo:
                # ---- elbo_global_local_ (:440-495): c = 1 re-derivations from the post-sweep sums (:456-467)
                eta_e = eta + n_eff / 2 + sums["colsum_gam"] / 2
                kappa_e = kappa + kappa_rate(sums, sig2_inv_vb) / 2
                nu_e = nu + sum_gam / 2
                rho_e = rho + tau_dot_m2 / 2
                log_tau_e = sp.digamma(eta_e) - np.log(kappa_e)
                log_sig2_inv_e = float(sp.digamma(nu_e) - math.log(rho_e))
                log_sig02_inv_vb = float(sp.digamma(nu_s0_vb) - math.log(rho_s0_vb))
                log_xi_inv_vb = float(sp.digamma(nu_xi_inv_vb) - math.log(rho_xi_inv_vb))
                # A: e_y_ (R/elbo.R:135-146)
                A = np.sum(-n_eff / 2 * math.log(2 * math.pi) + n_eff / 2 * log_tau_e
                           - tau_vb * (kappa_e - colsum_m2 * sig2_inv_vb / 2 - kappa))
                # B: e_beta_gamma_ (R/elbo.R:10-34): per-trait terms from the column sums + the device part
                if mis_pat is None:
                    gam_log_s2b = sums["colsum_gam"] * np.log(sig2_beta_vb)
                else:
                    gam_log_s2b = sums["colsum_gam_logsig2b"]  # sum_j gam log sig2_beta_vb(j,k) (R/elbo.R:28-30)
                B_loc = (np.sum(sums["colsum_gam"] * (log_sig2_inv_e / 2 + log_tau_e / 2 + 0.5) + gam_log_s2b / 2)
                         - np.sum(colsum_m2 * tau_vb) * sig2_inv_vb / 2 + elbo_b_dev - p * q * sig2_zeta_vb / 2
                         - q * np.sum(sig2_theta_vb) / 2)
                # D: e_zeta_ without its constants (R/elbo.R:153-161);  E: e_tau_ (:63-68)
                D_loc = -t02_inv * np.sum((zeta_vb - n0) ** 2) / 2
                E = np.sum((eta - eta_e) * log_tau_e - (kappa - kappa_e) * tau_vb + eta * np.log(kappa)
                           - eta_e * np.log(kappa_e) - sp.gammaln(eta) + sp.gammaln(eta_e))
                tot = comm.allreduce_sum(np.array([A + B_loc + D_loc + E, sum_zeta_local]))
                sum_zeta = float(tot[1])
                D_cst = (vec_sum_log_det_zeta - q_total * t02_inv * sig2_zeta_vb + q_total) / 2
                # C: e_theta_hs_, df = 1 (R/elbo.R:88-92)
                C = float(np.sum((log_sig02_inv_vb + math.log(shr_fac_inv)) / 2
                                 - sig02_inv_vb * shr_fac_inv * lam2_inv_vb
                                 * (theta_vb ** 2 + sig2_theta_vb - 2 * m0 * theta_vb + m0 ** 2) / 2
                                 + (np.log(sig2_theta_vb) + 1) / 2 - math.log(math.pi) + L_vb * lam2_inv_vb
                                 + np.log(Q_app)))
                F = (-0.5 * log_sig02_inv_vb - xi_inv_vb * sig02_inv_vb + log_xi_inv_vb / 2 - sp.gammaln(0.5)
                     - (nu_s0_vb - 1) * log_sig02_inv_vb + rho_s0_vb * sig02_inv_vb
                     - nu_s0_vb * math.log(rho_s0_vb) + sp.gammaln(nu_s0_vb))  # R/elbo.R:49-56
                G = _e_sig2_inv(0.5, nu_xi_inv_vb, log_xi_inv_vb, A2_inv, rho_xi_inv_vb, xi_inv_vb)
                H = _e_sig2_inv(nu, nu_e, log_sig2_inv_e, rho, rho_e, sig2_inv_vb)
                lb_new = float(tot[0] + D_cst + C + F + G + H)
                rec["lb"] = lb_new
                if verbose != 0 and comm.rank == 0 and (it == it_init or it % max(5, batch_conv) == 0):
                    print(f"ELBO = {lb_new}\n")
                if debug and lb_new + eps < lb_old:  # :359-360
                    raise RuntimeError("ELBO not increasing monotonically. Exit. ")
                diff_lb = abs(lb_new - lb_old)
                sum_exceed = int(np.sum(diff_lb > times_conv_sched * tol))  # :364
                if sum_exceed == 0:
                    converged = True
                elif ind_batch_conv > sum_exceed:
                    ind_batch_conv = sum_exceed
                    batch_conv = batch_conv_sched[ind_batch_conv - 1]
            else:
                sum_zeta = float(comm.allreduce_sum(np.array([sum_zeta_local]))[0])
            _lap("elbo")
            if trace is not None:
                rec["c_next"] = c
                rec["c"] = c_prev
                rec["host_ms"] = seg
                trace.append(rec)
            if not rec["annealing"]:  # the reference checkpoints in the non-annealed branch only (:338, :379-381)
                checkpoint_(it, checkpoint_path, ctx, theta_vb, zeta_vb, converged, lb_new, lb_old, lam2_inv_vb,
                            sig02_inv_vb, comm=comm, rate=checkpoint_rate,
                            extra=dict(tau_vb=tau_vb, sig2_beta_vb=sig2_beta_vb, sig2_theta_vb=sig2_theta_vb,
                                       has_missing=mis_pat is not None))

        checkpoint_join_(ctx)
        if not keep_checkpoints:
            checkpoint_clean_up_(checkpoint_path, comm)  # :388
        if iter_hook is not None:
            iter_hook(it + 1, ctx)
        if verbose != 0 and comm.rank == 0:
            if converged:
                print(f"Convergence obtained after {it} iterations. \nOptimal marginal log-likelihood variational "
                      f"lower bound (ELBO) = {lb_new}. \n")
            else:
                import warnings
                warnings.warn("Maximal number of iterations reached before convergence. Exit.")

        lb_opt = lb_new
        st = ctx.get_state(gam=True, mu=bool(full_output), beta=True)
        out = dict(beta_vb=st["beta_vb"], gam_vb=st["gam_vb"], theta_vb=theta_vb, zeta_vb=zeta_vb, n=n, p=p, q=q,
                   anneal=anneal, converged=converged, it=it, maxit=maxit, tol=tol, lb_opt=lb_opt,
                   diff_lb=abs(lb_opt - lb_old))  # :414-428
        if full_output:  # :404-410 (no Gram objects exist in sample space)
            out.update(mu_beta_vb=st["mu_beta_vb"], eta_vb=eta_vb, kappa_vb=kappa_vb, lam2_inv_vb=lam2_inv_vb,
                       nu_s0_vb=nu_s0_vb, nu_vb=nu_vb, nu_xi_inv_vb=nu_xi_inv_vb, rho_s0_vb=rho_s0_vb, rho_vb=rho_vb,
                       rho_xi_inv_vb=rho_xi_inv_vb, shr_fac_inv=shr_fac_inv, sig02_inv_vb=sig02_inv_vb,
                       sig2_beta_vb=sig2_beta_vb, sig2_inv_vb=sig2_inv_vb, sig2_theta_vb=sig2_theta_vb,
                       sig2_zeta_vb=sig2_zeta_vb, tau_vb=tau_vb, xi_inv_vb=xi_inv_vb, cp_Y_X=None, cp_X=None,
                       cp_X_Xbeta=None)
        return out
    finally:
        try:
            checkpoint_join_(ctx)
        except Exception:
            pass
        if own_ctx:
            ctx.close()
