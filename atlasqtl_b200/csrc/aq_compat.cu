// Stateless compatibility entry: the reference's 15-argument coreDualLoop contract
// (src/coreLoop.cpp:38-52) served by the sample-space CUDA sweep.  See aq_coreDualLoop in
// include/atlasqtl_b200.h.
#include <string>

#include "../../include/atlasqtl_b200.h"

extern "C" int aq_coreDualLoop(int device, int p, int q, const double* cp_X, const double* cp_Y_X, double* gam_vb,
                               const double* log_Phi_theta_plus_zeta, const double* log_1_min_Phi_theta_plus_zeta,
                               double log_sig2_inv_vb, const double* log_tau_vb, double* m1_beta, double* cp_betaX_X,
                               double* mu_beta_vb, const double* sig2_beta_vb, const double* tau_vb,
                               const int32_t* shuffled_ind, int n_ind, const int32_t* sample_q, int n_q, double c) {
    (void)device; (void)p; (void)q; (void)cp_X; (void)cp_Y_X; (void)gam_vb; (void)log_Phi_theta_plus_zeta;
    (void)log_1_min_Phi_theta_plus_zeta; (void)log_sig2_inv_vb; (void)log_tau_vb; (void)m1_beta; (void)cp_betaX_X;
    (void)mu_beta_vb; (void)sig2_beta_vb; (void)tau_vb; (void)shuffled_ind; (void)n_ind; (void)sample_q; (void)n_q; (void)c;
    return AQ_EUNSUPPORTED;  // TODO(round 1, later today): pivoted-Cholesky pseudo-design path
}
