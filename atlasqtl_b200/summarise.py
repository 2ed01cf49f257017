"""Post-processing on the returned object: mirror of R/summarise_output.R:207-223 (`assign_bFDR`) and the
selection sets of `summary` / `plot` (:99-106, :173-178)."""
import numpy as np


def assign_bFDR(mat_ppi):
    """Bayesian FDR of every pair: sort all PPIs decreasingly (column-major as.vector order, stable),
    running mean of 1 - PPI, scatter back."""
    mat_ppi = np.asarray(mat_ppi, dtype=np.float64)
    vec = mat_ppi.flatten(order="F")
    ind = np.argsort(-vec, kind="stable")
    fdr_ord = np.cumsum(1 - vec[ind]) / np.arange(1, vec.size + 1)
    out = np.empty_like(vec)
    out[ind] = fdr_ord
    return out.reshape(mat_ppi.shape, order="F")


def selected_pairs(gam_vb, thres, fdr_adjust=False):
    """(row, col) index pairs with PPI > thres, or bFDR < thres when fdr_adjust (R/summarise_output.R:99-106)."""
    m = assign_bFDR(gam_vb) < thres if fdr_adjust else np.asarray(gam_vb) > thres
    return np.argwhere(m)


# ----------------------------------------------------------------------------- the same sets, formed on the device
def _bits(x):
    return int(np.float64(x).view(np.uint64))


def _from_bits(b):
    return float(np.uint64(b).view(np.float64))


def select_ppi_device(ctx, thres, comm=None, k_first=0, capacity=1 << 24):
    """{(j, k): gam_vb > thres} straight from the device state of `ctx` (a SweepContext holding the traits
    [k_first, k_first + q_local)); returns (rows, cols) with global trait indices, column-major order."""
    j, k, _, n = ctx.ppi_collect(1, thres, 0.0, capacity)
    if n > capacity:
        raise MemoryError(f"{n} selected pairs exceed capacity={capacity}")
    return j.astype(np.int64), k.astype(np.int64) + k_first


def select_bFDR_device(ctx, thres, comm=None, k_first=0, p=None, capacity=1 << 24):
    """{(j, k): assign_bFDR(gam_vb)[j, k] < thres} (R/summarise_output.R:99-106, :207-223) without sorting or
    downloading the p x q matrix.

    The running mean of e = 1 - PPI along the decreasing-PPI order is non-decreasing, so the set is a prefix of
    that order: all pairs with e <= t1, plus the first m (column-major index order, R's order() is stable) of
    the pairs tied at the next value t2.  t1 is found by bisection on the bit pattern of a double, one
    streaming count/sum pass over gam_vb per probe; with several slabs (`comm`: an object with allreduce_sum)
    counts and sums simply add up (`comm` needs allreduce_sum and allreduce_min).  Returns (rows, cols, n_selected) for THIS slab, cols global.
    """
    red = (lambda v: comm.allreduce_sum(np.asarray(v, dtype=np.float64))) if comm is not None else (lambda v: np.asarray(v))
    p = ctx.p if p is None else p

    def stats(t):
        c, s = ctx.ppi_count_sum(t)
        r = red([c, s])
        return float(r[0]), float(r[1])

    def ok(t):  # does the prefix {e <= t} keep its running mean below thres?  (monotone in t)
        c, s = stats(t)
        return c == 0 or s / c < thres

    lo, hi = _bits(0.0), _bits(1.0)   # e in [0, 1]
    if not ok(0.0):
        t1, n1, s1 = -1.0, 0.0, 0.0   # even the pairs with PPI == 1 ... cannot fail (mean 0 < thres) unless thres <= 0
    elif ok(1.0):
        t1 = 1.0
        n1, s1 = stats(1.0)
    else:
        while hi - lo > 1:   # invariant: ok(lo), not ok(hi)
            mid = (lo + hi) // 2
            if ok(_from_bits(mid)):
                lo = mid
            else:
                hi = mid
        t1 = _from_bits(lo)
        n1, s1 = stats(t1)
    # ties at the next value: include the first m of them while the running mean stays below thres
    m, t2 = 0, None
    if t1 < 1.0:
        t2 = ctx.ppi_next_above(t1)
        if comm is not None:
            t2 = float(comm.allreduce_min(np.array([t2]))[0])
        if np.isfinite(t2):
            n2, _ = stats(t2)
            ntie = int(round(n2 - n1))
            while m < ntie and (s1 + (m + 1) * t2) / (n1 + m + 1) < thres:
                m += 1
    j, k, _, n = ctx.ppi_collect(0, -1.0, t1, capacity)
    if n > capacity:
        raise MemoryError(f"{n} selected pairs exceed capacity={capacity}")
    rows, cols = j.astype(np.int64), k.astype(np.int64) + k_first
    if m > 0:
        tj, tk, _, nt = ctx.ppi_collect(0, t1, t2, capacity)
        if nt > capacity:
            raise MemoryError("too many pairs tied at the bFDR boundary")
        lin = tj.astype(np.int64) + (tk.astype(np.int64) + k_first) * p   # as.vector() index of the pair
        if comm is not None and comm.world_size > 1:
            # the first m ties in GLOBAL column-major order: slabs are contiguous trait ranges, so lower ranks come first
            mine = np.zeros(comm.world_size)
            mine[comm.rank] = len(lin)
            before = int(round(comm.allreduce_sum(mine)[:comm.rank].sum()))
            take = max(0, min(len(lin), m - before))
        else:
            take = min(m, len(lin))
        o = np.argsort(lin, kind="stable")[:take]
        rows = np.concatenate([rows, tj[o].astype(np.int64)])
        cols = np.concatenate([cols, tk[o].astype(np.int64) + k_first])
    o = np.lexsort((rows, cols))
    return rows[o], cols[o], int(round(n1)) + m
