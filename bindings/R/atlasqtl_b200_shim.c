/*
 * .Call shim between R and libatlasqtl_b200.so (include/atlasqtl_b200.h).
 *
 * Replaces the generated RcppEigen glue of the reference (src/RcppExports.cpp:17-74) for the CAVI sweep:
 * plain R C API only (no Rcpp, no Eigen), one SEXP wrapper per C-ABI entry point, the context held in
 * an external pointer with a finalizer, errors raised with Rf_error() AFTER the C call has returned
 * (the library never throws and keeps no sticky CUDA error).
 *
 * Build inside the atlasqtl package (src/):   PKG_LIBS = -L<dir> -latlasqtl_b200
 * NOT compiled in this repository's image (no R headers there): it is deliberately logic-free so that
 * everything that can be tested is tested through the C ABI itself.
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
static double* dbl_or_null(SEXP x) { return Rf_isNull(x) ? NULL : REAL(x); }

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
  check(aq_set_order(get_ctx(ptr), Rf_isNull(shuffled_ind) ? NULL : (const int32_t*)INTEGER(shuffled_ind)));
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
  int q = Rf_ncols(gam_vb);
  double* o[5];
  SEXP res = PROTECT(sums_list(q, 0, o));
  check(aq_set_state(c, REAL(gam_vb), REAL(mu_beta_vb), o[0], o[1], o[2], o[3]));
  UNPROTECT(1);
  return res;
}

/* In place on caller-allocated p x q matrices, like coreDualLoop's in-place outputs (src/coreLoop.cpp:40,45-47) */
SEXP _atlasqtl_aq_get_state(SEXP ptr, SEXP gam_vb, SEXP mu_beta_vb, SEXP beta_vb) {
  check(aq_get_state(get_ctx(ptr), dbl_or_null(gam_vb), dbl_or_null(mu_beta_vb), dbl_or_null(beta_vb)));
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
  check(aq_snapshot_fetch(get_ctx(ptr), dbl_or_null(gam_vb), dbl_or_null(mu_beta_vb), dbl_or_null(beta_vb)));
  return R_NilValue;
}

SEXP _atlasqtl_aq_refresh_tables(SEXP ptr, SEXP theta_vb, SEXP zeta_vb, SEXP c_next, SEXP want_elbo) {
  double part = NA_REAL;
  check(aq_refresh_tables(get_ctx(ptr), REAL(theta_vb), REAL(zeta_vb), Rf_asReal(c_next),
                          Rf_asLogical(want_elbo) ? &part : NULL));
  return Rf_ScalarReal(part);
}

SEXP _atlasqtl_aq_sweep(SEXP ptr, SEXP c, SEXP log_sig2_inv_vb, SEXP tau_vb, SEXP log_tau_vb, SEXP sig2_beta_vb) {
  aq_ctx* ctx = get_ctx(ptr);
  int q = Rf_length(tau_vb);
  double* o[5];
  SEXP res = PROTECT(sums_list(q, 1, o));
  check(aq_sweep(ctx, Rf_asReal(c), Rf_asReal(log_sig2_inv_vb), REAL(tau_vb), REAL(log_tau_vb), REAL(sig2_beta_vb),
                 o[0], o[1], o[2], o[3], o[4]));
  UNPROTECT(1);
  return res;
}

SEXP _atlasqtl_aq_rowsums_zpart(SEXP ptr, SEXP p) {
  SEXP v = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(p)));
  check(aq_rowsums_zpart(get_ctx(ptr), REAL(v)));
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
  SEXP n_obs = PROTECT(Rf_allocVector(REALSXP, Rf_ncols(mis_pat)));
  check(aq_set_missing(get_ctx(ptr), REAL(mis_pat), REAL(n_obs)));
  UNPROTECT(1);
  return n_obs;
}

SEXP _atlasqtl_aq_set_state_mis(SEXP ptr, SEXP gam_vb, SEXP mu_beta_vb) {
  static const char* names[] = {"colsum_gam", "colsum_gam_mu2", "colsum_beta2", "resid_sq", "colsum_xn_gam",
                                "colsum_xn_gam_mu2", "colsum_xn_beta2"};
  double* o[7];
  SEXP res = PROTECT(named_qvecs(Rf_ncols(gam_vb), names, 7, o));
  check(aq_set_state_mis(get_ctx(ptr), REAL(gam_vb), REAL(mu_beta_vb), o[0], o[1], o[2], o[3], o[4], o[5], o[6]));
  UNPROTECT(1);
  return res;
}

SEXP _atlasqtl_aq_sweep_mis(SEXP ptr, SEXP c, SEXP log_sig2_inv_vb, SEXP sig2_inv_vb, SEXP tau_vb, SEXP log_tau_vb) {
  static const char* names[] = {"colsum_gam", "colsum_gam_mu2", "colsum_sig2b_gam", "colsum_xn_gam_mu2",
                                "colsum_xn_sig2b_gam", "colsum_xn_beta2", "resid_sq", "colsum_zpart", "colsum_gam_logsig2b"};
  double* o[9];
  SEXP res = PROTECT(named_qvecs(Rf_length(tau_vb), names, 9, o));
  check(aq_sweep_mis(get_ctx(ptr), Rf_asReal(c), Rf_asReal(log_sig2_inv_vb), Rf_asReal(sig2_inv_vb), REAL(tau_vb),
                     REAL(log_tau_vb), o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8]));
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
  int p = Rf_nrows(gam_vb), q = Rf_ncols(gam_vb);
  check(aq_coreDualLoop(0, p, q, REAL(cp_X), REAL(cp_Y_X), REAL(gam_vb), REAL(log_Phi), REAL(log_1_min_Phi),
                        Rf_asReal(log_sig2_inv_vb), REAL(log_tau_vb), REAL(m1_beta), REAL(cp_betaX_X), REAL(mu_beta_vb),
                        REAL(sig2_beta_vb), REAL(tau_vb), (const int32_t*)INTEGER(shuffled_ind), Rf_length(shuffled_ind),
                        (const int32_t*)INTEGER(sample_q), Rf_length(sample_q), Rf_asReal(c)));
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
    {NULL, NULL, 0}};

void R_init_atlasqtl(DllInfo* dll) { /* same registration as src/RcppExports.cpp:71-74 */
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
