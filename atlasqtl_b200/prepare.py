"""Pre-processing: host-side mirror of `prepare_data_` (reference R/prepare_atlasqtl.R:8-88).

One-off O(np) work that establishes the kernel's pre-conditions: columns of X centred and scaled with
the n-1 standard deviation (so X_j'X_j = n-1), constant and duplicated columns dropped, Y centred.
"""
import numpy as np


def rm_constant_(X_scaled, names):
    """R/utils.R:276-300: columns that became NaN after scale() were constant."""
    bool_cst = np.isnan(X_scaled).any(axis=0)
    rmvd = [names[j] for j in np.flatnonzero(bool_cst)]
    return X_scaled[:, ~bool_cst], bool_cst, rmvd


def rm_collinear_(X_scaled, names):
    """R/utils.R:303-343: drop exact duplicates (keep the first), remember who was dropped in favour of whom."""
    seen = {}
    bool_coll = np.zeros(X_scaled.shape[1], dtype=bool)
    rmvd = {}
    for j in range(X_scaled.shape[1]):
        key = X_scaled[:, j].tobytes()
        if key in seen:
            bool_coll[j] = True
            rmvd.setdefault(names[seen[key]], []).append(names[j])
        else:
            seen[key] = j
    return X_scaled[:, ~bool_coll], bool_coll, rmvd


def check_common_(Y, n, tol, maxit, checkpoint_path=None):
    """The argument checks shared by the host and device pre-processing (R/prepare_atlasqtl.R:11-45)."""
    import os
    if not (tol > 0):
        raise ValueError("tol must be positive.")
    if not (maxit >= 1 and int(maxit) == maxit):
        raise ValueError("maxit must be a natural number.")
    if checkpoint_path is not None and not os.path.isdir(os.path.dirname(checkpoint_path) or "."):  # :20-22 (path is a prefix)
        raise ValueError("The directory specified in checkpoint_path does not exist. Please make sure to provide a "
                         "valid path.")
    if Y.shape[0] != n:
        raise ValueError("X and Y must have the same number of samples.")
    obs = ~np.isnan(Y)
    if obs.sum() / Y.size < 0.05:  # :39-40
        raise ValueError("Too few non-NA values in matrix Y. Exit.")
    if (obs.sum(axis=0) / n < 0.025).any():  # :42-45
        raise ValueError("Column(s) of matrix Y have more than 97.5% missing values, and should be removed. Exit.")


def prepare_data_(Y, X, tol, maxit, user_seed=None, verbose=0, checkpoint_path=None):
    """Returns dict(Y, X, bool_rmvd_x, initial_colnames_X, rmvd_cst_x, rmvd_coll_x, names_x, names_y)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    if X.ndim != 2 or Y.ndim != 2:
        raise ValueError("X and Y must be matrices.")
    n, p = X.shape
    if n < 2:
        raise ValueError("X must have at least 2 rows.")  # check_structure_/check dims, :11-30
    check_common_(Y, n, tol, maxit, checkpoint_path)
    if np.isnan(X).any():
        raise ValueError("X must not contain missing values.")
    names_x = [f"Cov_x_{j + 1}" for j in range(p)]  # :54
    names_y = [f"Resp_{k + 1}" for k in range(Y.shape[1])]  # :55
    with np.errstate(invalid="ignore", divide="ignore"):
        Xs = (X - X.mean(axis=0)) / X.std(axis=0, ddof=1)  # scale(X), :57
    Xs, bool_cst, rmvd_cst = rm_constant_(Xs, names_x)
    names_after_cst = [nm for nm, b in zip(names_x, bool_cst) if not b]
    Xs, bool_coll, rmvd_coll = rm_collinear_(Xs, names_after_cst)
    bool_rmvd = bool_cst.copy()
    bool_rmvd[~bool_cst] = bool_coll  # :68-69
    if Xs.shape[1] < 1:
        raise ValueError("There must be at least 1 non-constant candidate predictor stored in X.")
    Yc = Y - np.nanmean(Y, axis=0)  # scale(Y, center = TRUE, scale = FALSE), :83
    kept = [nm for nm, b in zip(names_after_cst, bool_coll) if not b]
    return dict(Y=np.asfortranarray(Yc), X=np.asfortranarray(Xs), bool_rmvd_x=bool_rmvd,
                initial_colnames_X=names_after_cst, rmvd_cst_x=rmvd_cst, rmvd_coll_x=rmvd_coll, names_x=kept,
                names_y=names_y)


def prepare_data_device_(Y, X, tol, maxit, user_seed=None, verbose=0, checkpoint_path=None, *, packed_n=None, device=0):
    """`prepare_data_` with the O(np) work on the GPU (include/atlasqtl_b200.h, aq_prep_x / aq_prep_geno): same
    checks, same returned fields, except that "X" is a `PreparedPredictors` handle (the standardised matrix exists
    only on the device) and "Y" is the RAW response matrix, centred on the device when the context is created.
    X is the raw n x p matrix, or -- with packed_n = n -- packed genotype calls (`device.pack_genotypes`)."""
    from .device import PreparedPredictors
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim != 2:
        raise ValueError("X and Y must be matrices.")
    n = int(packed_n) if packed_n is not None else np.shape(X)[0]
    check_common_(Y, n, tol, maxit, checkpoint_path)
    prep = PreparedPredictors(packed=X, n=n, device=device) if packed_n is not None else PreparedPredictors(X, device=device)
    if prep.p < 1:
        raise ValueError("There must be at least 1 non-constant candidate predictor stored in X.")
    names_x = [f"Cov_x_{j + 1}" for j in range(prep.p_raw)]
    names_y = [f"Resp_{k + 1}" for k in range(Y.shape[1])]
    rmvd_cst = [names_x[j] for j in np.flatnonzero(prep.status == 1)]
    rmvd_coll = {}
    for j in np.flatnonzero(prep.status == 2):
        rmvd_coll.setdefault(names_x[prep.dup_of[j]], []).append(names_x[j])
    return dict(Y=np.asfortranarray(Y), X=prep, bool_rmvd_x=prep.status != 0,
                initial_colnames_X=[names_x[j] for j in np.flatnonzero(prep.status != 1)], rmvd_cst_x=rmvd_cst,
                rmvd_coll_x=rmvd_coll, names_x=[names_x[j] for j in np.flatnonzero(prep.status == 0)], names_y=names_y)
