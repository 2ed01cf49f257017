/*
 * .Call shim between R and libatlasqtl_b200.so (include/atlasqtl_b200.h).
 *
 * Replaces the generated RcppEigen glue of the reference (src/RcppExports.cpp:17-74) for the CAVI sweep:
 * plain R C API only (no Rcpp, no Eigen), one SEXP wrapper per C-ABI entry point, the context held in
 * an external pointer with a finalizer, errors raised with Rf_error() AFTER the C call has returned
 * (the library never throws and keeps no sticky CUDA error).
 *
 * Build inside the atlasqtl package (src/):   PKG_LIBS = -L<dir> -latlasqtl_b200
 * NOT compiled against R in this repository's image (no R headers there; tests/test_cabi.py only syntax-checks it against
 * a stand-in for the few R API declarations it uses): it is deliberately logic-free so that everything that can be
 * tested is tested through the C ABI itself.  What it does own is the R-side contract: every SEXP is checked for type
 * and for the size the CONTEXT expects (aq_dims) before a pointer is handed to the library, so a wrong-sized argument
 * from R is an R error, never an out-of-bounds access.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "atlasqtl_b200.h"

static void ctx_finalizer(SEXP ptr) {
  aq_ctx* c = (aq_ctx*)R_ExternalPtrAddr(ptr);
  if (c) { aq_destroy(c); R_ClearExternalPtr(ptr); }
}
static aq_ctx* get_ctx(SEXP ptr) {
  aq_ctx* c = (aq_ctx*)R_ExternalPtrAddr(ptr);
  if (!c) Rf_error("atlasqtl_b200: context already destroyed");
  return c;
}
static void check(int rc) { if (rc != AQ_OK) Rf_error("atlasqtl_b200 (%d): %s", rc, aq_last_error()); }

/* ---- argument validation against the context's dimensions */
typedef struct { int n, p, q; } dims_t;
static dims_t ctx_dims(aq_ctx* c) {
  dims_t d;
  check(aq_dims(c, &d.n, &d.p, &d.q, NULL, NULL));
  return d;
}
static double* need_mat(SEXP x, const char* name, int nrow, int ncol) {
  if (!Rf_isReal(x) || !Rf_isMatrix(x)) Rf_error("%s must be a double matrix", name);
  if (Rf_nrows(x) != nrow || Rf_ncols(x) != ncol) Rf_error("%s must be %d x %d (is %d x %d)", name, nrow, ncol, Rf_nrows(x), Rf_ncols(x));
  return REAL(x);
}
static double* need_mat_or_null(SEXP x, const char* name, int nrow, int ncol) {
  return Rf_isNull(x) ? NULL : need_mat(x, name, nrow, ncol);
}
static double* need_vec(SEXP x, const char* name, R_xlen_t len) {
  if (!Rf_isReal(x)) Rf_error("%s must be a double vector", name);
  if (XLENGTH(x) != len) Rf_error("%s must have length %ld (has %ld)", name, (long)len, (long)XLENGTH(x));
  return REAL(x);
}
static const int32_t* need_ind(SEXP x, const char* name, R_xlen_t len) {
  if (!Rf_isInteger(x)) Rf_error("%s must be an integer vector (0-based, as.integer())", name);
  if (len >= 0 && XLENGTH(x) != len) Rf_error("%s must have length %ld (has %ld)", name, (long)len, (long)XLENGTH(x));
  return (const int32_t*)INTEGER(x);
}
static double need_scalar(SEXP x, const char* name) {
  if (!(Rf_isReal(x) || Rf_isInteger(x)) || XLENGTH(x) != 1) Rf_error("%s must be a single number", name);
  return Rf_asReal(x);
}

/* aq_create(X, Y, device): X n x p, Y n x q double matrices (storage mode checked like Eigen::Map does) */
SEXP _atlasqtl_aq_create(SEXP X, SEXP Y, SEXP device) {
  if (!Rf_isReal(X) || !Rf_isReal(Y) || !Rf_isMatrix(X) || !Rf_isMatrix(Y)) Rf_error("X and Y must be double matrices");
  int n = Rf_nrows(X), p = Rf_ncols(X), q = Rf_ncols(Y);
  if (Rf_nrows(Y) != n) Rf_error("X and Y must have the same number of rows");
  aq_ctx* c = NULL;
  check(aq_create(&c, Rf_asInteger(device), n, p, q, REAL(X), REAL(Y)));
  SEXP ptr = PROTECT(R_MakeExternalPtr(c, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}
SEXP _atlasqtl_aq_destroy(SEXP ptr) { ctx_finalizer(ptr); return R_NilValue; }

/* shuffled_ind: integer vector, 0-based (as.integer(0:(p-1)) in the reference) or NULL */
SEXP _atlasqtl_aq_set_order(SEXP ptr, SEXP shuffled_ind) {
  aq_ctx* c = get_ctx(ptr);
  check(aq_set_order(c, Rf_isNull(shuffled_ind) ? NULL : need_ind(shuffled_ind, "shuffled_ind", ctx_dims(c).p)));
  return R_NilValue;
}

static SEXP sums_list(int q, int with_z, double** out) {
  const char* names[] = {"colsum_gam", "colsum_gam_mu2", "colsum_beta2", "resid_sq", "colsum_zpart"};
  int k = with_z ? 5 : 4;
  SEXP res = PROTECT(Rf_allocVector(VECSXP, k)), nm = PROTECT(Rf_allocVector(STRSXP, k));
  for (int i = 0; i < k; ++i) {
    SEXP v = PROTECT(Rf_allocVector(REALSXP, q));
    out[i] = REAL(v);
    SET_VECTOR_ELT(res, i, v);
    SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
    UNPROTECT(1);
  }
  Rf_setAttrib(res, R_NamesSymbol, nm);
  UNPROTECT(2);
  return res;
}

SEXP _atlasqtl_aq_set_state(SEXP ptr, SEXP gam_vb, SEXP mu_beta_vb) {
  aq_ctx* c = get_ctx(ptr);
  dims_t d = ctx_dims(c);
  const double* g = need_mat(gam_vb, "gam_vb", d.p, d.q);
  const double* m = need_mat(mu_beta_vb, "mu_beta_vb", d.p, d.q);
  double* o[5];
  SEXP res = PROTECT(sums_list(d.q, 0, o));
  check(aq_set_state(c, g, m, o[0], o[1], o[2], o[3]));
  UNPROTECT(1);
  return res;
}

/* In place on caller-allocated p x q matrices, like coreDualLoop's in-place outputs (src/coreLoop.cpp:40,45-47) */
SEXP _atlasqtl_aq_get_state(SEXP ptr, SEXP gam_vb, SEXP mu_beta_vb, SEXP beta_vb) {
  aq_ctx* c = get_ctx(ptr);
  dims_t d = ctx_dims(c);
  check(aq_get_state(c, need_mat_or_null(gam_vb, "gam_vb", d.p, d.q), need_mat_or_null(mu_beta_vb, "mu_beta_vb", d.p, d.q),
                     need_mat_or_null(beta_vb, "beta_vb", d.p, d.q)));
  return R_NilValue;
}

/* Checkpoints without a stall: freeze the state on the device now (milliseconds) ... */
SEXP _atlasqtl_aq_snapshot(SEXP ptr) {
  check(aq_snapshot(get_ctx(ptr)));
  return R_NilValue;
}
/* ... and fill the matrices later (e.g. just before saveRDS in checkpoint_): the sweeps issued in between ran on the main
 * stream while nothing was copied; a package that owns a worker thread may call aq_snapshot_fetch from it instead */
SEXP _atlasqtl_aq_snapshot_fetch(SEXP ptr, SEXP gam_vb, SEXP mu_beta_vb, SEXP beta_vb) {
  aq_ctx* c = get_ctx(ptr);
  dims_t d = ctx_dims(c);
  check(aq_snapshot_fetch(c, need_mat_or_null(gam_vb, "gam_vb", d.p, d.q), need_mat_or_null(mu_beta_vb, "mu_beta_vb", d.p, d.q),
                          need_mat_or_null(beta_vb, "beta_vb", d.p, d.q)));
  return R_NilValue;
}

SEXP _atlasqtl_aq_refresh_tables(SEXP ptr, SEXP theta_vb, SEXP zeta_vb, SEXP c_next, SEXP want_elbo) {
  double part = NA_REAL;
  aq_ctx* c = get_ctx(ptr);
  dims_t d = ctx_dims(c);
  check(aq_refresh_tables(c, need_vec(theta_vb, "theta_vb", d.p), need_vec(zeta_vb, "zeta_vb", d.q),
                          need_scalar(c_next, "c_next"), Rf_asLogical(want_elbo) == TRUE ? &part : NULL));
  return Rf_ScalarReal(part);
}

SEXP _atlasqtl_aq_sweep(SEXP ptr, SEXP c, SEXP log_sig2_inv_vb, SEXP tau_vb, SEXP log_tau_vb, SEXP sig2_beta_vb) {
  aq_ctx* ctx = get_ctx(ptr);
  dims_t d = ctx_dims(ctx);
  const double* tau = need_vec(tau_vb, "tau_vb", d.q);
  const double* ltau = need_vec(log_tau_vb, "log_tau_vb", d.q);
  const double* s2 = need_vec(sig2_beta_vb, "sig2_beta_vb", d.q);
  double* o[5];
  SEXP res = PROTECT(sums_list(d.q, 1, o));
  check(aq_sweep(ctx, need_scalar(c, "c"), need_scalar(log_sig2_inv_vb, "log_sig2_inv_vb"), tau, ltau, s2,
                 o[0], o[1], o[2], o[3], o[4]));
  UNPROTECT(1);
  return res;
}

SEXP _atlasqtl_aq_rowsums_zpart(SEXP ptr, SEXP p) {   /* p is kept for the call signature; the length is the context's */
  aq_ctx* c = get_ctx(ptr);
  dims_t d = ctx_dims(c);
  if (!Rf_isNull(p) && Rf_asInteger(p) != d.p) Rf_error("p does not match the context (%d)", d.p);
  SEXP v = PROTECT(Rf_allocVector(REALSXP, d.p));
  check(aq_rowsums_zpart(c, REAL(v)));
  UNPROTECT(1);
  return v;
}

/* ---- missing responses: replaces `_atlasqtl_coreDualMisLoop` (src/RcppExports.cpp:41-63) and its set-up
 * (X_norm_sq, cp_X_rm; R/atlasqtl_global_local_core.R:19-33) */
static SEXP named_qvecs(int q, const char** names, int k, double** out) {
  SEXP res = PROTECT(Rf_allocVector(VECSXP, k)), nm = PROTECT(Rf_allocVector(STRSXP, k));
  for (int i = 0; i < k; ++i) {
    SEXP v = PROTECT(Rf_allocVector(REALSXP, q));
    out[i] = REAL(v);
    SET_VECTOR_ELT(res, i, v);
    SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
    UNPROTECT(1);
  }
  Rf_setAttrib(res, R_NamesSymbol, nm);
  UNPROTECT(2);
  return res;
}

/* mis_pat: the n x q double matrix ifelse(is.na(Y), 0, 1); returns colSums(mis_pat) */
SEXP _atlasqtl_aq_set_missing(SEXP ptr, SEXP mis_pat) {
  aq_ctx* c = get_ctx(ptr);
  dims_t d = ctx_dims(c);
  const double* mp = need_mat(mis_pat, "mis_pat", d.n, d.q);
  SEXP n_obs = PROTECT(Rf_allocVector(REALSXP, d.q));
  check(aq_set_missing(c, mp, REAL(n_obs)));
  UNPROTECT(1);
  return n_obs;
}

SEXP _atlasqtl_aq_set_state_mis(SEXP ptr, SEXP gam_vb, SEXP mu_beta_vb) {
  static const char* names[] = {"colsum_gam", "colsum_gam_mu2", "colsum_beta2", "resid_sq", "colsum_xn_gam",
                                "colsum_xn_gam_mu2", "colsum_xn_beta2"};
  double* o[7];
  aq_ctx* c = get_ctx(ptr);
  dims_t d = ctx_dims(c);
  const double* g = need_mat(gam_vb, "gam_vb", d.p, d.q);
  const double* m = need_mat(mu_beta_vb, "mu_beta_vb", d.p, d.q);
  SEXP res = PROTECT(named_qvecs(d.q, names, 7, o));
  check(aq_set_state_mis(c, g, m, o[0], o[1], o[2], o[3], o[4], o[5], o[6]));
  UNPROTECT(1);
  return res;
}

SEXP _atlasqtl_aq_sweep_mis(SEXP ptr, SEXP c, SEXP log_sig2_inv_vb, SEXP sig2_inv_vb, SEXP tau_vb, SEXP log_tau_vb) {
  static const char* names[] = {"colsum_gam", "colsum_gam_mu2", "colsum_sig2b_gam", "colsum_xn_gam_mu2",
                                "colsum_xn_sig2b_gam", "colsum_xn_beta2", "resid_sq", "colsum_zpart", "colsum_gam_logsig2b"};
  double* o[9];
  aq_ctx* ctx = get_ctx(ptr);
  dims_t d = ctx_dims(ctx);
  const double* tau = need_vec(tau_vb, "tau_vb", d.q);
  const double* ltau = need_vec(log_tau_vb, "log_tau_vb", d.q);
  SEXP res = PROTECT(named_qvecs(d.q, names, 9, o));
  check(aq_sweep_mis(ctx, need_scalar(c, "c"), need_scalar(log_sig2_inv_vb, "log_sig2_inv_vb"),
                     need_scalar(sig2_inv_vb, "sig2_inv_vb"), tau, ltau, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8]));
  UNPROTECT(1);
  return res;
}

/* prepare_data_ on the device (R/prepare_atlasqtl.R:57-83): X raw n x p double matrix, or a raw vector of packed
 * 2-bit calls (bytes_per_col * p bytes) with n given.  Returns list(prep = <external pointer>, status, dup_of, mean, sd). */
static void prep_finalizer(SEXP ptr) {
  aq_prep* P = (aq_prep*)R_ExternalPtrAddr(ptr);
  if (P) { aq_prep_destroy(P); R_ClearExternalPtr(ptr); }
}
static SEXP prep_result(aq_prep* P, int p_raw) {
  static const char* names[] = {"prep", "status", "dup_of", "mean", "sd"};
  SEXP res = PROTECT(Rf_allocVector(VECSXP, 5)), nm = PROTECT(Rf_allocVector(STRSXP, 5));
  SEXP ptr = PROTECT(R_MakeExternalPtr(P, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, prep_finalizer, TRUE);
  SEXP status = PROTECT(Rf_allocVector(RAWSXP, p_raw)), dup = PROTECT(Rf_allocVector(INTSXP, p_raw));
  SEXP mean = PROTECT(Rf_allocVector(REALSXP, p_raw)), sd = PROTECT(Rf_allocVector(REALSXP, p_raw));
  int rc = aq_prep_result(P, RAW(status), (int32_t*)INTEGER(dup), REAL(mean), REAL(sd));
  SET_VECTOR_ELT(res, 0, ptr); SET_VECTOR_ELT(res, 1, status); SET_VECTOR_ELT(res, 2, dup);
  SET_VECTOR_ELT(res, 3, mean); SET_VECTOR_ELT(res, 4, sd);
  for (int i = 0; i < 5; ++i) SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
  Rf_setAttrib(res, R_NamesSymbol, nm);
  UNPROTECT(7);
  check(rc);
  return res;
}
SEXP _atlasqtl_aq_prep_x(SEXP X, SEXP device) {
  if (!Rf_isReal(X) || !Rf_isMatrix(X)) Rf_error("X must be a double matrix");
  aq_prep* P = NULL;
  check(aq_prep_x(&P, Rf_asInteger(device), Rf_nrows(X), Rf_ncols(X), REAL(X), NULL));
  return prep_result(P, Rf_ncols(X));
}
SEXP _atlasqtl_aq_prep_geno(SEXP geno, SEXP n, SEXP p, SEXP bytes_per_col, SEXP device) {
  if (TYPEOF(geno) != RAWSXP) Rf_error("geno must be a raw vector");
  int pp = Rf_asInteger(p);
  double bpc = Rf_asReal(bytes_per_col);
  if ((double)XLENGTH(geno) < bpc * pp) Rf_error("geno is shorter than bytes_per_col * p");
  aq_prep* P = NULL;
  check(aq_prep_geno(&P, Rf_asInteger(device), Rf_asInteger(n), pp, RAW(geno), (int64_t)bpc, NULL));
  return prep_result(P, pp);
}
/* context over the kept columns; Y raw (NA = missing).  Returns the context with attribute "n_obs" = colSums(!is.na(Y)) */
SEXP _atlasqtl_aq_create_prepared(SEXP prep, SEXP Y) {
  aq_prep* P = (aq_prep*)R_ExternalPtrAddr(prep);
  if (!P) Rf_error("atlasqtl_b200: prep already destroyed");
  if (!Rf_isReal(Y) || !Rf_isMatrix(Y)) Rf_error("Y must be a double matrix");
  int n_prep = 0;
  check(aq_prep_dims(P, &n_prep, NULL, NULL));
  if (Rf_nrows(Y) != n_prep) Rf_error("X and Y must have the same number of samples.");
  SEXP n_obs = PROTECT(Rf_allocVector(REALSXP, Rf_ncols(Y)));
  aq_ctx* c = NULL;
  int rc = aq_create_prepared(&c, P, Rf_ncols(Y), REAL(Y), REAL(n_obs));   /* NA_real_ is a NaN: read as missing */
  if (rc != AQ_OK) { UNPROTECT(1); check(rc); }
  SEXP ptr = PROTECT(R_MakeExternalPtr(c, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
  Rf_setAttrib(ptr, Rf_install("n_obs"), n_obs);
  UNPROTECT(2);
  return ptr;
}

/* Stateless drop-in with the reference's exact 15 arguments (src/RcppExports.cpp:17). */
SEXP _atlasqtl_coreDualLoop(SEXP cp_X, SEXP cp_Y_X, SEXP gam_vb, SEXP log_Phi, SEXP log_1_min_Phi, SEXP log_sig2_inv_vb,
                            SEXP log_tau_vb, SEXP m1_beta, SEXP cp_betaX_X, SEXP mu_beta_vb, SEXP sig2_beta_vb,
                            SEXP tau_vb, SEXP shuffled_ind, SEXP sample_q, SEXP c) {
  if (!Rf_isReal(gam_vb) || !Rf_isMatrix(gam_vb)) Rf_error("gam_vb must be a double matrix");
  int p = Rf_nrows(gam_vb), q = Rf_ncols(gam_vb);
  check(aq_coreDualLoop(0, p, q, need_mat(cp_X, "cp_X", p, p), need_mat(cp_Y_X, "cp_Y_X", q, p), REAL(gam_vb),
                        need_mat(log_Phi, "log_Phi_theta_plus_zeta", p, q),
                        need_mat(log_1_min_Phi, "log_1_min_Phi_theta_plus_zeta", p, q),
                        need_scalar(log_sig2_inv_vb, "log_sig2_inv_vb"), need_vec(log_tau_vb, "log_tau_vb", q),
                        need_mat(m1_beta, "m1_beta", p, q), need_mat(cp_betaX_X, "cp_betaX_X", p, q),
                        need_mat(mu_beta_vb, "mu_beta_vb", p, q), need_vec(sig2_beta_vb, "sig2_beta_vb", q),
                        need_vec(tau_vb, "tau_vb", q), need_ind(shuffled_ind, "shuffled_ind", -1), Rf_length(shuffled_ind),
                        need_ind(sample_q, "sample_q", -1), Rf_length(sample_q), need_scalar(c, "c")));
  return R_NilValue;
}

/* Stateless drop-in with the reference's exact 16 arguments (src/RcppExports.cpp:41): cp_X_rm is the list of q p x p
 * matrices built at R/atlasqtl_global_local_core.R:25-32, sig2_beta_vb is p x q. */
SEXP _atlasqtl_coreDualMisLoop(SEXP cp_X, SEXP cp_X_rm, SEXP cp_Y_X, SEXP gam_vb, SEXP log_Phi, SEXP log_1_min_Phi,
                               SEXP log_sig2_inv_vb, SEXP log_tau_vb, SEXP m1_beta, SEXP cp_betaX_X, SEXP mu_beta_vb,
                               SEXP sig2_beta_vb, SEXP tau_vb, SEXP shuffled_ind, SEXP sample_q, SEXP c) {
  if (!Rf_isReal(gam_vb) || !Rf_isMatrix(gam_vb)) Rf_error("gam_vb must be a double matrix");
  int p = Rf_nrows(gam_vb), q = Rf_ncols(gam_vb);
  if (TYPEOF(cp_X_rm) != VECSXP || XLENGTH(cp_X_rm) != q) Rf_error("cp_X_rm must be a list of %d matrices", q);
  const double** rm = (const double**)R_alloc((size_t)q, sizeof(double*));   /* freed by R at the end of .Call */
  for (int k = 0; k < q; ++k) rm[k] = need_mat(VECTOR_ELT(cp_X_rm, k), "cp_X_rm[[k]]", p, p);
  check(aq_coreDualMisLoop(0, p, q, need_mat(cp_X, "cp_X", p, p), rm, need_mat(cp_Y_X, "cp_Y_X", q, p), REAL(gam_vb),
                           need_mat(log_Phi, "log_Phi_theta_plus_zeta", p, q),
                           need_mat(log_1_min_Phi, "log_1_min_Phi_theta_plus_zeta", p, q),
                           need_scalar(log_sig2_inv_vb, "log_sig2_inv_vb"), need_vec(log_tau_vb, "log_tau_vb", q),
                           need_mat(m1_beta, "m1_beta", p, q), need_mat(cp_betaX_X, "cp_betaX_X", p, q),
                           need_mat(mu_beta_vb, "mu_beta_vb", p, q), need_mat(sig2_beta_vb, "sig2_beta_vb", p, q),
                           need_vec(tau_vb, "tau_vb", q), need_ind(shuffled_ind, "shuffled_ind", -1),
                           Rf_length(shuffled_ind), need_ind(sample_q, "sample_q", -1), Rf_length(sample_q),
                           need_scalar(c, "c")));
  return R_NilValue;
}

static const R_CallMethodDef CallEntries[] = {
    {"_atlasqtl_aq_create", (DL_FUNC)&_atlasqtl_aq_create, 3},
    {"_atlasqtl_aq_destroy", (DL_FUNC)&_atlasqtl_aq_destroy, 1},
    {"_atlasqtl_aq_set_order", (DL_FUNC)&_atlasqtl_aq_set_order, 2},
    {"_atlasqtl_aq_set_state", (DL_FUNC)&_atlasqtl_aq_set_state, 3},
    {"_atlasqtl_aq_get_state", (DL_FUNC)&_atlasqtl_aq_get_state, 4},
    {"_atlasqtl_aq_snapshot", (DL_FUNC)&_atlasqtl_aq_snapshot, 1},
    {"_atlasqtl_aq_snapshot_fetch", (DL_FUNC)&_atlasqtl_aq_snapshot_fetch, 4},
    {"_atlasqtl_aq_refresh_tables", (DL_FUNC)&_atlasqtl_aq_refresh_tables, 5},
    {"_atlasqtl_aq_sweep", (DL_FUNC)&_atlasqtl_aq_sweep, 6},
    {"_atlasqtl_aq_rowsums_zpart", (DL_FUNC)&_atlasqtl_aq_rowsums_zpart, 2},
    {"_atlasqtl_aq_set_missing", (DL_FUNC)&_atlasqtl_aq_set_missing, 2},
    {"_atlasqtl_aq_set_state_mis", (DL_FUNC)&_atlasqtl_aq_set_state_mis, 3},
    {"_atlasqtl_aq_sweep_mis", (DL_FUNC)&_atlasqtl_aq_sweep_mis, 6},
    {"_atlasqtl_aq_prep_x", (DL_FUNC)&_atlasqtl_aq_prep_x, 2},
    {"_atlasqtl_aq_prep_geno", (DL_FUNC)&_atlasqtl_aq_prep_geno, 5},
    {"_atlasqtl_aq_create_prepared", (DL_FUNC)&_atlasqtl_aq_create_prepared, 2},
    {"_atlasqtl_coreDualLoop", (DL_FUNC)&_atlasqtl_coreDualLoop, 15},
    {"_atlasqtl_coreDualMisLoop", (DL_FUNC)&_atlasqtl_coreDualMisLoop, 16},
    {NULL, NULL, 0}};

void R_init_atlasqtl(DllInfo* dll) { /* same registration as src/RcppExports.cpp:71-74 */
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
