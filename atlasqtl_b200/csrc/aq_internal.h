// Library-internal entry points shared between translation units (not part of the C ABI).
#pragma once
#include "../../include/atlasqtl_b200.h"

namespace aq {
// Upload gam_vb / mu_beta_vb (p x q_local, column-major) WITHOUT rebuilding the residual: the residual is
// taken to be the Y handed to aq_create.  Used by the stateless compatibility entry, whose caller supplies
// the running cross-products instead of Y.
int internal_load_state(aq_ctx* ctx, const double* gam_vb, const double* mu_beta_vb);
// Upload D = log(1 - Phi) - log(Phi) directly (p x q_local, column-major); W and I0 are zeroed.
int internal_load_dtab(aq_ctx* ctx, const double* d_host);
// Upload an explicit p x q_local sig2_beta_vb (column-major) for the missing-response sweep, which otherwise forms
// 1 / (c (X_norm_sq + sig2_inv_vb) tau_vb) on the fly (R/update_vb.R:47): the stateless coreDualMisLoop entry takes
// sig2_beta_vb as an argument.
int internal_load_sig2(aq_ctx* ctx, const double* sig2_host);
// Record an error message for aq_last_error() and return `code`.
int internal_fail(int code, const char* msg);
}  // namespace aq
