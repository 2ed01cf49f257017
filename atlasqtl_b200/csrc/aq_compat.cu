// Stateless compatibility entries: the reference's 15-argument coreDualLoop contract (src/coreLoop.cpp:38-52; .Call glue
// src/RcppExports.cpp:17-38) and its 16-argument coreDualMisLoop contract (src/coreLoop.cpp:91-106; glue
// src/RcppExports.cpp:41-63) served by the sample-space CUDA sweeps.
//
// The reference hands over Gram quantities only (cp_X = X'X, cp_Y_X = Y'X, cp_betaX_X = X'X beta).  Any
// X~ with X~'X~ = cp_X reproduces every statistic the loop forms, so:
//   1. pivoted Cholesky  P' cp_X P = L L'  (rank r <= n - 1), X~ = L'  (r "pseudo-samples" x p);
//   2. residual R~ (r x q) from the pivot rows:  L_1 R~ = (cp_Y_X' - cp_betaX_X)[pivots, ]  (forward solve), so that
//      X~' R~ = X'Y - X'X beta, i.e. exactly the running cross-products the caller passed in;
//   3. one CUDA sweep on (X~, R~) with beta_old = m1_beta and D = log_1_min_Phi - log_Phi;
//   4. outputs in place: gam_vb, mu_beta_vb, m1_beta = gam * mu, cp_betaX_X = cp_Y_X' - L R~_new.
// With missing responses every trait k has its own Gram matrix cp_X - cp_X_rm[[k]] = L_k L_k': the pseudo-samples of the
// traits of a chunk are stacked into one design and trait k observes only its own block of rows (the masked-residual
// kernel then forms exactly X~_k' X~_k for it).
// Cost of 1./2./4. is O(p^2 r + p r q) on the host per call (per trait with missing responses), so these are parity /
// compatibility entries for the sizes where the reference itself can run (its p x p inputs exist), not the fast path.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "aq_internal.h"

using aq::internal_fail;

namespace {
// Pivoted Cholesky of the symmetric PSD p x p matrix G (column-major):  G[piv, piv] = L L',  L stored column-major
// p x rmax in pivoted row order, rank *r_out (pivots below 1e-13 max diag are taken as zero).
int pivoted_cholesky(const double* G, int p, int rmax_in, const char* who, std::vector<double>& L, std::vector<int>& piv,
                     int* r_out) {
    const size_t P = (size_t)p;
    std::vector<double> diag(p);
    piv.resize(p);
    double dmax0 = 0.0;
    for (int j = 0; j < p; ++j) {
        diag[j] = G[j + j * P];
        piv[j] = j;
        dmax0 = std::max(dmax0, diag[j]);
    }
    if (!(dmax0 > 0.0)) return internal_fail(AQ_EINVAL, (std::string(who) + ": the Gram matrix has no positive diagonal entry").c_str());
    const double tol = 1e-13 * dmax0;
    const int rmax = std::min(p, rmax_in);
    L.assign(P * rmax, 0.0);
    int r = 0;
    for (; r < rmax; ++r) {
        int best = r;
        for (int i = r + 1; i < p; ++i)
            if (diag[piv[i]] > diag[piv[best]]) best = i;
        if (diag[piv[best]] <= tol) break;
        std::swap(piv[r], piv[best]);
        for (int k = 0; k < r; ++k) std::swap(L[r + k * P], L[best + k * P]);
        const int jr = piv[r];
        const double lrr = std::sqrt(diag[jr]);
        L[r + r * P] = lrr;
        for (int i = r + 1; i < p; ++i) {
            const int ji = piv[i];
            double v = G[ji + jr * P];
            for (int k = 0; k < r; ++k) v -= L[i + k * P] * L[r + k * P];
            v /= lrr;
            L[i + r * P] = v;
            diag[ji] -= v * v;
        }
    }
    if (r == rmax && r < p) {
        double rest = 0.0;
        for (int i = r; i < p; ++i) rest = std::max(rest, diag[piv[i]]);
        if (rest > tol) return internal_fail(AQ_EUNSUPPORTED, (std::string(who) + ": rank of the Gram matrix exceeds " + std::to_string(rmax)).c_str());
    }
    if (r < 2) return internal_fail(AQ_EUNSUPPORTED, (std::string(who) + ": rank of the Gram matrix below 2").c_str());
    *r_out = r;
    return AQ_OK;
}
}  // namespace

extern "C" int aq_coreDualLoop(int device, int p, int q, const double* cp_X, const double* cp_Y_X, double* gam_vb,
                               const double* log_Phi_theta_plus_zeta, const double* log_1_min_Phi_theta_plus_zeta,
                               double log_sig2_inv_vb, const double* log_tau_vb, double* m1_beta, double* cp_betaX_X,
                               double* mu_beta_vb, const double* sig2_beta_vb, const double* tau_vb,
                               const int32_t* shuffled_ind, int n_ind, const int32_t* sample_q, int n_q, double c) {
    if (!cp_X || !cp_Y_X || !gam_vb || !log_Phi_theta_plus_zeta || !log_1_min_Phi_theta_plus_zeta || !log_tau_vb ||
        !m1_beta || !cp_betaX_X || !mu_beta_vb || !sig2_beta_vb || !tau_vb || !shuffled_ind || !sample_q)
        return internal_fail(AQ_EINVAL, "aq_coreDualLoop: NULL argument");
    if (p < 1 || q < 1 || n_q < 0 || n_q > q) return internal_fail(AQ_EINVAL, "aq_coreDualLoop: bad dimensions");
    if (n_ind != p) return internal_fail(AQ_EUNSUPPORTED, "aq_coreDualLoop: shuffled_ind must visit every SNP once");
    if (n_q == 0) return AQ_OK;
    std::vector<char> seen(q, 0);
    for (int a = 0; a < n_q; ++a) {
        if (sample_q[a] < 0 || sample_q[a] >= q || seen[sample_q[a]])
            return internal_fail(AQ_EINVAL, "aq_coreDualLoop: sample_q must hold distinct trait indices in 0..q-1");
        seen[sample_q[a]] = 1;
    }
    const size_t P = (size_t)p;
    // ---- 1. pivoted Cholesky of cp_X
    std::vector<double> L;
    std::vector<int> piv;
    int r = 0;
    int rc0 = pivoted_cholesky(cp_X, p, 1008, "aq_coreDualLoop", L, piv, &r);
    if (rc0 != AQ_OK) return rc0;
    // ---- X~ = L' in ORIGINAL SNP order: column j of X~ (r values) = row of L whose pivot is j
    std::vector<double> Xt((size_t)r * p);
    for (int i = 0; i < p; ++i)
        for (int k = 0; k < r; ++k) Xt[k + (size_t)piv[i] * r] = L[i + k * P];
    // ---- 2. residual from the caller's running cross-products (forward solve with the r x r pivot block)
    std::vector<double> Rt((size_t)r * n_q), gs(P * n_q), ms(P * n_q), ds(P * n_q);
    std::vector<double> tau_s(n_q), ltau_s(n_q), sig2_s(n_q);
    for (int a = 0; a < n_q; ++a) {
        const int k = sample_q[a];
        double* col = Rt.data() + (size_t)a * r;
        for (int i = 0; i < r; ++i) {
            const int j = piv[i];
            double v = cp_Y_X[k + (size_t)j * q] - cp_betaX_X[j + (size_t)k * P];
            for (int t = 0; t < i; ++t) v -= L[i + t * P] * col[t];
            col[i] = v / L[i + i * P];
        }
        for (int j = 0; j < p; ++j) {
            gs[j + (size_t)a * P] = 1.0;  // the kernel forms beta_old = gam_old * mu_old: carry m1_beta exactly
            ms[j + (size_t)a * P] = m1_beta[j + (size_t)k * P];
            ds[j + (size_t)a * P] = log_1_min_Phi_theta_plus_zeta[j + (size_t)k * P] - log_Phi_theta_plus_zeta[j + (size_t)k * P];
        }
        tau_s[a] = tau_vb[k];
        ltau_s[a] = log_tau_vb[k];
        sig2_s[a] = sig2_beta_vb[k];
    }
    // ---- 3. the CUDA sweep
    aq_ctx* ctx = nullptr;
    int rc = aq_create(&ctx, device, r, p, n_q, Xt.data(), Rt.data());
    if (rc != AQ_OK) return rc;
    rc = aq_set_order(ctx, shuffled_ind);
    if (rc == AQ_OK) rc = aq::internal_load_state(ctx, gs.data(), ms.data());
    if (rc == AQ_OK) rc = aq::internal_load_dtab(ctx, ds.data());
    if (rc == AQ_OK)
        rc = aq_sweep(ctx, c, log_sig2_inv_vb, tau_s.data(), ltau_s.data(), sig2_s.data(), nullptr, nullptr, nullptr,
                      nullptr, nullptr);
    if (rc == AQ_OK) rc = aq_get_state(ctx, gs.data(), ms.data(), ds.data());  // ds <- beta
    if (rc == AQ_OK) rc = aq_get_residual(ctx, Rt.data());
    aq_destroy(ctx);
    if (rc != AQ_OK) return rc;
    // ---- 4. outputs, in place, only for the swept traits
    for (int a = 0; a < n_q; ++a) {
        const int k = sample_q[a];
        const double* col = Rt.data() + (size_t)a * r;
        for (int j = 0; j < p; ++j) {
            gam_vb[j + (size_t)k * P] = gs[j + (size_t)a * P];
            mu_beta_vb[j + (size_t)k * P] = ms[j + (size_t)a * P];
            m1_beta[j + (size_t)k * P] = ds[j + (size_t)a * P];
        }
        for (int i = 0; i < p; ++i) {  // cp_betaX_X = X'Y - X~' R~
            double v = 0.0;
            for (int t = 0; t < r; ++t) v += L[i + t * P] * col[t];
            const int j = piv[i];
            cp_betaX_X[j + (size_t)k * P] = cp_Y_X[k + (size_t)j * q] - v;
        }
    }
    return AQ_OK;
}

// coreDualMisLoop (src/coreLoop.cpp:91-138): cp_X_rm[k] is the p x p matrix crossprod(X[missing rows of trait k, ])
// (R/atlasqtl_global_local_core.R:25-32), sig2_beta_vb is p x q.
extern "C" int aq_coreDualMisLoop(int device, int p, int q, const double* cp_X, const double* const* cp_X_rm,
                                  const double* cp_Y_X, double* gam_vb, const double* log_Phi_theta_plus_zeta,
                                  const double* log_1_min_Phi_theta_plus_zeta, double log_sig2_inv_vb,
                                  const double* log_tau_vb, double* m1_beta, double* cp_betaX_X, double* mu_beta_vb,
                                  const double* sig2_beta_vb, const double* tau_vb, const int32_t* shuffled_ind, int n_ind,
                                  const int32_t* sample_q, int n_q, double c) {
    if (!cp_X || !cp_X_rm || !cp_Y_X || !gam_vb || !log_Phi_theta_plus_zeta || !log_1_min_Phi_theta_plus_zeta ||
        !log_tau_vb || !m1_beta || !cp_betaX_X || !mu_beta_vb || !sig2_beta_vb || !tau_vb || !shuffled_ind || !sample_q)
        return internal_fail(AQ_EINVAL, "aq_coreDualMisLoop: NULL argument");
    if (p < 1 || q < 1 || n_q < 0 || n_q > q) return internal_fail(AQ_EINVAL, "aq_coreDualMisLoop: bad dimensions");
    if (n_ind != p) return internal_fail(AQ_EUNSUPPORTED, "aq_coreDualMisLoop: shuffled_ind must visit every SNP once");
    if (n_q == 0) return AQ_OK;
    std::vector<char> seen(q, 0);
    for (int a = 0; a < n_q; ++a) {
        if (sample_q[a] < 0 || sample_q[a] >= q || seen[sample_q[a]])
            return internal_fail(AQ_EINVAL, "aq_coreDualMisLoop: sample_q must hold distinct trait indices in 0..q-1");
        seen[sample_q[a]] = 1;
        if (!cp_X_rm[sample_q[a]]) return internal_fail(AQ_EINVAL, "aq_coreDualMisLoop: NULL element in cp_X_rm");
    }
    const size_t P = (size_t)p;
    constexpr int kMaxRows = 2048;   // sample capacity of the masked-residual kernel
    struct Fac {
        std::vector<double> L;
        std::vector<int> piv;
        int r = 0;
    };
    std::vector<double> Gk(P * P);
    int a0 = 0;
    while (a0 < n_q) {
        // ---- a chunk of traits whose stacked pseudo-samples fit one context
        std::vector<Fac> fac;
        int rows = 0, a1 = a0;
        for (; a1 < n_q; ++a1) {
            const int k = sample_q[a1];
            const double* rm = cp_X_rm[k];
            for (size_t i = 0; i < P * P; ++i) Gk[i] = cp_X[i] - rm[i];   // :120, :132
            Fac f;
            int rc = pivoted_cholesky(Gk.data(), p, kMaxRows, "aq_coreDualMisLoop", f.L, f.piv, &f.r);
            if (rc != AQ_OK) return rc;
            if (rows + f.r > kMaxRows) break;
            rows += f.r;
            fac.push_back(std::move(f));
        }
        const int nq = a1 - a0;
        if (nq == 0) return internal_fail(AQ_EUNSUPPORTED, "aq_coreDualMisLoop: Gram rank exceeds the kernel's sample capacity");
        const size_t N = (size_t)rows;
        std::vector<double> Xt(N * p, 0.0), Rt(N * nq, 0.0), mis(N * nq, 0.0);
        std::vector<double> gs(P * nq), ms(P * nq), ds(P * nq), s2(P * nq), tau_s(nq), ltau_s(nq);
        int row0 = 0;
        std::vector<int> first(nq);
        for (int a = 0; a < nq; ++a) {
            const int k = sample_q[a0 + a];
            const Fac& f = fac[a];
            first[a] = row0;
            for (int i = 0; i < p; ++i)   // X~_k = L_k' in original SNP order, in the rows of this trait
                for (int t = 0; t < f.r; ++t) Xt[row0 + t + (size_t)f.piv[i] * N] = f.L[i + t * P];
            double* col = Rt.data() + (size_t)a * N + row0;
            for (int i = 0; i < f.r; ++i) {   // L_k1 R~_k = (X'Y - X'X beta)[pivots, k]
                const int j = f.piv[i];
                double v = cp_Y_X[k + (size_t)j * q] - cp_betaX_X[j + (size_t)k * P];
                for (int t = 0; t < i; ++t) v -= f.L[i + t * P] * col[t];
                col[i] = v / f.L[i + i * P];
                mis[(size_t)a * N + row0 + i] = 1.0;
            }
            for (int j = 0; j < p; ++j) {
                gs[j + (size_t)a * P] = 1.0;
                ms[j + (size_t)a * P] = m1_beta[j + (size_t)k * P];
                ds[j + (size_t)a * P] = log_1_min_Phi_theta_plus_zeta[j + (size_t)k * P] - log_Phi_theta_plus_zeta[j + (size_t)k * P];
                s2[j + (size_t)a * P] = sig2_beta_vb[j + (size_t)k * P];
            }
            tau_s[a] = tau_vb[k];
            ltau_s[a] = log_tau_vb[k];
            row0 += f.r;
        }
        aq_ctx* ctx = nullptr;
        int rc = aq_create(&ctx, device, rows, p, nq, Xt.data(), Rt.data());
        if (rc != AQ_OK) return rc;
        rc = aq_set_order(ctx, shuffled_ind);
        if (rc == AQ_OK) rc = aq_set_missing(ctx, mis.data(), nullptr);   // X_norm_sq(j,k) = diag of trait k's Gram matrix
        if (rc == AQ_OK) rc = aq::internal_load_state(ctx, gs.data(), ms.data());
        if (rc == AQ_OK) rc = aq::internal_load_dtab(ctx, ds.data());
        if (rc == AQ_OK) rc = aq::internal_load_sig2(ctx, s2.data());
        if (rc == AQ_OK)
            rc = aq_sweep_mis(ctx, c, log_sig2_inv_vb, 1.0, tau_s.data(), ltau_s.data(), nullptr, nullptr, nullptr, nullptr,
                              nullptr, nullptr, nullptr, nullptr, nullptr);
        if (rc == AQ_OK) rc = aq_get_state(ctx, gs.data(), ms.data(), ds.data());  // ds <- beta
        if (rc == AQ_OK) rc = aq_get_residual(ctx, Rt.data());
        aq_destroy(ctx);
        if (rc != AQ_OK) return rc;
        for (int a = 0; a < nq; ++a) {
            const int k = sample_q[a0 + a];
            const Fac& f = fac[a];
            const double* col = Rt.data() + (size_t)a * N + first[a];
            for (int j = 0; j < p; ++j) {
                gam_vb[j + (size_t)k * P] = gs[j + (size_t)a * P];
                mu_beta_vb[j + (size_t)k * P] = ms[j + (size_t)a * P];
                m1_beta[j + (size_t)k * P] = ds[j + (size_t)a * P];
            }
            for (int i = 0; i < p; ++i) {   // cp_betaX_X[, k] = X'Y[, k] - X~_k' R~_k
                double v = 0.0;
                for (int t = 0; t < f.r; ++t) v += f.L[i + t * P] * col[t];
                const int j = f.piv[i];
                cp_betaX_X[j + (size_t)k * P] = cp_Y_X[k + (size_t)j * q] - v;
            }
        }
        a0 = a1;
    }
    return AQ_OK;
}
