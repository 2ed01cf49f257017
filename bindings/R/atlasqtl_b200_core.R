# Call-site patch for atlasqtl_global_local_core_ (reference R/atlasqtl_global_local_core.R).
# Only the lines that touch p x q / n x q objects change; every p-, q- and scalar-sized update_*_vb_
# helper of R/update_vb.R and every e_*_ term of R/elbo.R is called exactly as before.
# (atlasqtl_b200/core.py is the same patch written in Python, and is what the tests in this repository run.)
#
#   reference lines                         replacement
#   :40-42   Y_norm_sq, cp_X, cp_Y_X    ->  ctx <- .Call(`_atlasqtl_aq_create`, X, Y, 0L)
#   :61-63   log_Phi tables             ->  .Call(`_atlasqtl_aq_refresh_tables`, ctx, theta_vb, zeta_vb, c, FALSE)
#   :112-115 beta_vb, m2_beta, cp_X_Xbeta -> s <- .Call(`_atlasqtl_aq_set_state`, ctx, gam_vb, mu_beta_vb)
#   :134-135 sum(gam_vb), colSums(m2_beta) -> sum(s$colsum_gam); s$colsum_gam_mu2 + sig2_beta_vb * s$colsum_gam
#   :141-142 eta_vb, kappa_vb           ->  c * (eta + n/2 + s$colsum_gam/2) - c + 1
#                                           c * (kappa + (s$resid_sq + (n-1+sig2_inv_vb) * colsum_m2 - (n-1) * s$colsum_beta2)/2)
#   :162-170 coreDualLoop(...)          ->  s <- .Call(`_atlasqtl_aq_sweep`, ctx, c, log_sig2_inv_vb, tau_vb, log_tau_vb, sig2_beta_vb)
#   :235-237 m2_beta, Z                 ->  rowsums_Z <- .Call(`_atlasqtl_aq_rowsums_zpart`, ctx, p) / sqrt_c + q * theta_vb + sum(zeta_vb)
#                                           colsums_Z <- s$colsum_zpart / sqrt_c + sum(theta_vb) + p * zeta_vb
#   :280,290 rowSums(Z), colSums(Z)     ->  rowsums_Z, colsums_Z in update_theta_vb_ / update_zeta_vb_
#   :293-295 log_Phi tables             ->  elbo_B_dev <- .Call(`_atlasqtl_aq_refresh_tables`, ctx, theta_vb, zeta_vb, c_next, want_elbo)
#   :472     e_beta_gamma_(...)         ->  sum(s$colsum_gam * (log_sig2_inv_vb/2 + log_tau_vb/2 + (log(sig2_beta_vb)+1)/2)) -
#                                           sum(colsum_m2 * tau_vb) * sig2_inv_vb/2 + elbo_B_dev - p*q*sig2_zeta_vb/2 - q*sum(sig2_theta_vb)/2
#   :418-428 output                     ->  gam_vb <- beta_vb <- matrix(0, p, q); .Call(`_atlasqtl_aq_get_state`, ctx, gam_vb, NULL, beta_vb)
#
# Missing responses (any(is.na(Y)), :19-38 and the mis_pat branches of R/update_vb.R / R/elbo.R):
#   :21-33   mis_pat, X_norm_sq, cp_X_rm  ->  mis_pat <- ifelse(is.na(Y), 0, 1); Y[is.na(Y)] <- 0
#                                             n_obs <- .Call(`_atlasqtl_aq_set_missing`, ctx, mis_pat)      # = colSums(mis_pat)
#   :112-115 (m2_beta with sweep = TRUE)   ->  s <- .Call(`_atlasqtl_aq_set_state_mis`, ctx, gam_vb, mu_beta_vb)
#                                             colsum_m2 <- s$colsum_gam_mu2 + sig2_beta_vb * s$colsum_gam
#                                             colsum_xn_m2 <- s$colsum_xn_gam_mu2 + sig2_beta_vb * s$colsum_xn_gam
#   :141     update_eta_vb_(.., mis_pat)   ->  c * (eta + n_obs/2 + s$colsum_gam/2) - c + 1
#   :142     update_kappa_vb_(.., X_norm_sq) -> c * (kappa + (s$resid_sq + sig2_inv_vb * colsum_m2 + colsum_xn_m2 - s$colsum_xn_beta2)/2)
#   :147     update_sig2_beta_vb_(.., X_norm_sq) -> formed on the device inside the sweep (p x q never exists in R)
#   :172-175 coreDualMisLoop(...)          ->  s <- .Call(`_atlasqtl_aq_sweep_mis`, ctx, c, log_sig2_inv_vb, sig2_inv_vb, tau_vb, log_tau_vb)
#   :235     m2_beta (p x q sig2_beta_vb)  ->  colsum_m2 <- s$colsum_gam_mu2 + s$colsum_sig2b_gam
#                                             colsum_xn_m2 <- s$colsum_xn_gam_mu2 + s$colsum_xn_sig2b_gam
#   :470     e_y_(.., mis_pat)             ->  arg <- n_obs * (log_tau_vb - log(2 * pi)) / 2
#   :472     e_beta_gamma_ (p x q sig2_beta_vb) -> sum(s$colsum_gam * (log_sig2_inv_vb/2 + log_tau_vb/2 + 1/2) + s$colsum_gam_logsig2b/2) - ...

coreDualLoop <- function(cp_X, cp_Y_X, gam_vb, log_Phi_theta_plus_zeta, log_1_min_Phi_theta_plus_zeta,
                         log_sig2_inv_vb, log_tau_vb, m1_beta, cp_betaX_X, mu_beta_vb, sig2_beta_vb, tau_vb,
                         shuffled_ind, sample_q, c = 1) {
  # unchanged closure (R/RcppExports.R:4-6): the symbol now resolves to the CUDA-backed shim
  invisible(.Call(`_atlasqtl_coreDualLoop`, cp_X, cp_Y_X, gam_vb, log_Phi_theta_plus_zeta,
                  log_1_min_Phi_theta_plus_zeta, log_sig2_inv_vb, log_tau_vb, m1_beta, cp_betaX_X, mu_beta_vb,
                  sig2_beta_vb, tau_vb, shuffled_ind, sample_q, c))
}
#
# Pre-processing on the device (prepare_data_, R/prepare_atlasqtl.R:57-83; optional, see INTEGRATION.md section 2e):
#   :57-72   scale(X), rm_constant_, rm_collinear_  ->  pr <- .Call(`_atlasqtl_aq_prep_x`, X, 0L)   # or _atlasqtl_aq_prep_geno
#                                                       bool_cst_x <- pr$status == 1L; bool_rmvd_x <- pr$status != 0L
#   :83      scale(Y, center = TRUE, scale = FALSE) ->  ctx <- .Call(`_atlasqtl_aq_create_prepared`, pr$prep, Y)  (replaces aq_create)
