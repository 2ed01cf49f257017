/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * extern "C" entry points around the reference's OWN coreDualLoop /
 * coreDualMisLoop (/root/reference/src/coreLoop.cpp:38-86, :91-138), which the
 * Makefile compiles unmodified from where it lies, against oracle/shim/RcppEigen.h.
 * Argument order follows the generated .Call glue (src/RcppExports.cpp:17-38).
 */
#include "utils.h"   /* the reference's header: pulls the shim + atlasqtl_types.h */

void coreDualLoop(const MapMat cp_X, const MapMat cp_Y_X, MapArr2D gam_vb,
                  const MapArr2D log_Phi_theta_plus_zeta, const MapArr2D log_1_min_Phi_theta_plus_zeta,
                  const double log_sig2_inv_vb, const MapArr1D log_tau_vb, MapMat m1_beta,
                  MapMat cp_betaX_X, MapArr2D mu_beta_vb, const MapArr1D sig2_beta_vb,
                  const MapArr1D tau_vb, const Eigen::VectorXi shuffled_ind,
                  const Eigen::VectorXi sample_q, const double c);
void coreDualMisLoop(const MapMat cp_X, const List cp_X_rm, const MapMat cp_Y_X, MapArr2D gam_vb,
                     const MapArr2D log_Phi_theta_plus_zeta, const MapArr2D log_1_min_Phi_theta_plus_zeta,
                     const double log_sig2_inv_vb, const MapArr1D log_tau_vb, MapMat m1_beta,
                     MapMat cp_betaX_X, MapArr2D mu_beta_vb, const MapArr2D sig2_beta_vb,
                     const MapArr1D tau_vb, const Eigen::VectorXi shuffled_ind,
                     const Eigen::VectorXi sample_q, const double c);

extern "C" {

void ref_coreDualLoop(int p, int q, double* cp_X, double* cp_Y_X, double* gam_vb, double* log_Phi,
                      double* log_1_min_Phi, double log_sig2_inv_vb, double* log_tau_vb,
                      double* m1_beta, double* cp_betaX_X, double* mu_beta_vb, double* sig2_beta_vb,
                      double* tau_vb, const int* shuffled_ind, int n_ind, const int* sample_q,
                      int n_q, double c) {
  coreDualLoop(MapMat(cp_X, p, p), MapMat(cp_Y_X, q, p), MapArr2D(gam_vb, p, q),
               MapArr2D(log_Phi, p, q), MapArr2D(log_1_min_Phi, p, q), log_sig2_inv_vb,
               MapArr1D(log_tau_vb, q), MapMat(m1_beta, p, q), MapMat(cp_betaX_X, p, q),
               MapArr2D(mu_beta_vb, p, q), MapArr1D(sig2_beta_vb, q), MapArr1D(tau_vb, q),
               Eigen::VectorXi(shuffled_ind, n_ind), Eigen::VectorXi(sample_q, n_q), c);
}

/* cp_X_rm: q consecutive p-by-p column-major matrices. sig2_beta_vb: p-by-q. */
void ref_coreDualMisLoop(int p, int q, double* cp_X, double* cp_X_rm, double* cp_Y_X,
                         double* gam_vb, double* log_Phi, double* log_1_min_Phi,
                         double log_sig2_inv_vb, double* log_tau_vb, double* m1_beta,
                         double* cp_betaX_X, double* mu_beta_vb, double* sig2_beta_vb,
                         double* tau_vb, const int* shuffled_ind, int n_ind, const int* sample_q,
                         int n_q, double c) {
  List lst;
  for (int k = 0; k < q; ++k) lst.push_back(cp_X_rm + (std::size_t)k * p * p, p, p);
  coreDualMisLoop(MapMat(cp_X, p, p), lst, MapMat(cp_Y_X, q, p), MapArr2D(gam_vb, p, q),
                  MapArr2D(log_Phi, p, q), MapArr2D(log_1_min_Phi, p, q), log_sig2_inv_vb,
                  MapArr1D(log_tau_vb, q), MapMat(m1_beta, p, q), MapMat(cp_betaX_X, p, q),
                  MapArr2D(mu_beta_vb, p, q), MapArr2D(sig2_beta_vb, p, q), MapArr1D(tau_vb, q),
                  Eigen::VectorXi(shuffled_ind, n_ind), Eigen::VectorXi(sample_q, n_q), c);
}

}  /* extern "C" */
