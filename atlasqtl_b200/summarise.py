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
