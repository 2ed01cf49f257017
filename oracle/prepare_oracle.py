"""TEST INFRASTRUCTURE -- CPU restatement of the reference's pre-processing, `prepare_data_`
(/root/reference R/prepare_atlasqtl.R:57-83) with `rm_constant_` / `rm_collinear_` (R/utils.R:276-343).

Only tests/ may import this module; the product never does.  Parity unpinned by the reference's own tests (it has
none for this step and no R exists here to run it): what pins it is R's documented semantics, restated literally:

  scale(X)            centre = colMeans(X) (R accumulates in long double: restated with exactly rounded sums, math.fsum),
                      then x - centre, then divide by sqrt(sum(x_centred^2) / (n - 1))      (base R `scale.default`)
  rm_constant_        bool_cst = is.nan(colSums(mat)): a constant column became 0/0 = NaN   (R/utils.R:278)
  rm_collinear_       bool_coll = duplicated(mat, MARGIN = 2): a column equal, value by value, to an EARLIER column;
                      names(rmvd_coll) = the kept column it equals                           (R/utils.R:305, :327-333)
  scale(Y, center = TRUE, scale = FALSE)   colMeans(Y, na.rm = TRUE) subtracted, NA stays NA (R/prepare_atlasqtl.R:83)

Pure Python loops over columns: meant for test sizes.
"""
import math

import numpy as np


def scale_(X):
    X = np.asarray(X, dtype=np.float64)
    n, p = X.shape
    out = np.empty((n, p), order="F")
    for j in range(p):
        col = X[:, j]
        centre = math.fsum(col) / n
        d = col - centre
        sd = math.sqrt(math.fsum(d * d) / (n - 1))
        with np.errstate(invalid="ignore", divide="ignore"):
            out[:, j] = d / sd
    return out


def prepare_data_(Y, X):
    """Returns dict(X, Y, bool_cst_x, bool_coll_x (over the non-constant columns), bool_rmvd_x, kept (0-based raw
    indices), dup_of (raw index of the kept twin for every removed duplicate, -1 elsewhere))."""
    Xs = scale_(X)
    p = Xs.shape[1]
    bool_cst = np.isnan(Xs.sum(axis=0))                       # rm_constant_
    idx = np.flatnonzero(~bool_cst)
    seen = {}
    bool_coll = np.zeros(idx.size, dtype=bool)
    dup_of = np.full(p, -1, dtype=np.int64)
    for pos, j in enumerate(idx):                             # duplicated(mat, MARGIN = 2)
        key = tuple(Xs[:, j])                                 # value equality (0.0 == -0.0, like identical(num.eq = TRUE))
        if key in seen:
            bool_coll[pos] = True
            dup_of[j] = seen[key]
        else:
            seen[key] = j
    bool_rmvd = bool_cst.copy()
    bool_rmvd[~bool_cst] = bool_coll                          # R/prepare_atlasqtl.R:68-69
    kept = np.flatnonzero(~bool_rmvd)
    Y = np.asarray(Y, dtype=np.float64)
    Yc = np.empty(Y.shape, order="F")
    for k in range(Y.shape[1]):
        obs = ~np.isnan(Y[:, k])
        Yc[:, k] = Y[:, k] - math.fsum(Y[obs, k]) / int(obs.sum())
    return dict(X=np.asfortranarray(Xs[:, kept]), Y=Yc, bool_cst_x=bool_cst, bool_coll_x=bool_coll, bool_rmvd_x=bool_rmvd,
                kept=kept, dup_of=dup_of)
