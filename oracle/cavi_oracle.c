/*
 * TEST INFRASTRUCTURE ONLY -- the CPU oracle for the atlasqtl CAVI sweep.
 * Nothing in the product path (atlasqtl_b200/) may link, load or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, as the checker or as the timed CPU baseline.
 *
 * Three restatements of the same per-pair update (reference: src/coreLoop.cpp:56-85;
 * second statement of the formulas in R: R/atlasqtl_global_local_core.R:190-202):
 *
 *   oracle_core_dual_loop   -- dual / Gram form, trait-outer, SNP-inner, p-long column
 *                              axpy: the loop nest of src/coreLoop.cpp:58-85 on raw pointers.
 *   oracle_sweep_primal     -- the same arithmetic in sample space:
 *                              residual R = Y - X beta, s = X_j'r_k + beta_jk |X_j|^2,
 *                              r_k -= delta X_j.  Traits are independent, so the trait loop
 *                              is split over POSIX threads ("best-effort CPU").
 *   oracle_sweep_primal_blocked -- App. A3: blocks of B SNPs, S = X_B' R, exact in-block
 *                              Gauss-Seidel from the Gram block, one rank-B residual update.
 *                              This is the formulation the CUDA kernel implements.
 *
 * Parity status: the dual loop is pinned against the reference's own coreLoop.cpp compiled
 * unmodified (oracle/_ref, see Makefile) in tests/test_oracle.py.
 *
 * Layout everywhere: column-major (R layout).  X is n x p, Y/R n x q, all p x q arrays p x q.
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* src/coreLoop.cpp:28-33 */
static double log_one_plus_exp(double x) {
  double m = x;
  if (x < 0) m = 0;
  return log(exp(x - m) + exp(-m)) + m;
}

/* One (SNP j, trait k) update given the leave-one-out statistic s = X_j'(y_k - X beta_k + X_j beta_jk).
 * src/coreLoop.cpp:73-79.  Returns the new m1_beta; writes mu and gam. */
static inline double pair_update(double s, double c, double sig2_beta_k, double tau_k, double d_jk,
                                 double cst_k, double* mu_out, double* gam_out) {
  double mu = c * sig2_beta_k * tau_k * s;
  double gam = exp(-log_one_plus_exp(c * (d_jk - mu * mu / (2 * sig2_beta_k) + cst_k)));
  *mu_out = mu;
  *gam_out = gam;
  return gam * mu;
}

/* src/coreLoop.cpp:38-86.  cp_X p x p, cp_Y_X q x p, others p x q; in place on gam_vb, m1_beta,
 * cp_betaX_X, mu_beta_vb. */
void oracle_core_dual_loop(int p, int q, const double* cp_X, const double* cp_Y_X, double* gam_vb,
                           const double* log_Phi, const double* log_1_min_Phi, double log_sig2_inv_vb,
                           const double* log_tau_vb, double* m1_beta, double* cp_betaX_X,
                           double* mu_beta_vb, const double* sig2_beta_vb, const double* tau_vb,
                           const int* shuffled_ind, int n_ind, const int* sample_q, int n_q, double c) {
  double* cst = (double*)malloc(sizeof(double) * (size_t)q);
  for (int k = 0; k < q; ++k) /* :56 */
    cst[k] = -(log_tau_vb[k] + log_sig2_inv_vb + log(sig2_beta_vb[k])) / 2;
  for (int a = 0; a < n_q; ++a) {
    int k = sample_q[a];
    double* cbx = cp_betaX_X + (size_t)k * p;
    for (int b = 0; b < n_ind; ++b) {
      int j = shuffled_ind[b];
      size_t jk = (size_t)j + (size_t)k * p;
      double m1_old = m1_beta[jk];
      const double* gcol = cp_X + (size_t)j * p;
      double loo = cbx[j] - m1_old * gcol[j];                      /* :71 */
      double s = cp_Y_X[(size_t)k + (size_t)j * q] - loo;          /* :73 */
      double m1_new = pair_update(s, c, sig2_beta_vb[k], tau_vb[k],
                                  log_1_min_Phi[jk] - log_Phi[jk], cst[k],
                                  &mu_beta_vb[jk], &gam_vb[jk]);
      m1_beta[jk] = m1_new;
      double delta = m1_new - m1_old;
      for (int i = 0; i < p; ++i) cbx[i] += delta * gcol[i];       /* :81 */
    }
  }
  free(cst);
}

/* Primal sweep over the trait range [k0, k1).  X n x p (columns contiguous), xnorm2[j] = X_j'X_j,
 * R n x q residual (in/out), perm = shuffled_ind.  In place on gam_vb, m1_beta, mu_beta_vb, R. */
typedef struct {
  int n, p, q, k0, k1, n_ind;
  const double *X, *xnorm2, *log_Phi, *log_1_min_Phi, *log_tau_vb, *sig2_beta_vb, *tau_vb;
  double *R, *gam_vb, *m1_beta, *mu_beta_vb;
  double log_sig2_inv_vb, c;
  const int* perm;
} primal_args;

static void* primal_range(void* vp) {
  primal_args* a = (primal_args*)vp;
  int n = a->n, p = a->p;
  for (int k = a->k0; k < a->k1; ++k) {
    double cst = -(a->log_tau_vb[k] + a->log_sig2_inv_vb + log(a->sig2_beta_vb[k])) / 2;
    double* r = a->R + (size_t)k * n;
    for (int b = 0; b < a->n_ind; ++b) {
      int j = a->perm[b];
      size_t jk = (size_t)j + (size_t)k * p;
      const double* x = a->X + (size_t)j * n;
      double dot = 0;
      for (int i = 0; i < n; ++i) dot += x[i] * r[i];
      double m1_old = a->m1_beta[jk];
      double s = dot + m1_old * a->xnorm2[j];
      double m1_new = pair_update(s, a->c, a->sig2_beta_vb[k], a->tau_vb[k],
                                  a->log_1_min_Phi[jk] - a->log_Phi[jk], cst,
                                  &a->mu_beta_vb[jk], &a->gam_vb[jk]);
      a->m1_beta[jk] = m1_new;
      double delta = m1_new - m1_old;
      for (int i = 0; i < n; ++i) r[i] -= delta * x[i];
    }
  }
  return NULL;
}

/* Traits are independent inside a sweep, so they are split over `nthreads` POSIX threads
 * ("best-effort CPU"); nthreads <= 1 runs inline. */
void oracle_sweep_primal(int n, int p, int q, const double* X, const double* xnorm2, double* R,
                         double* gam_vb, const double* log_Phi, const double* log_1_min_Phi,
                         double log_sig2_inv_vb, const double* log_tau_vb, double* m1_beta,
                         double* mu_beta_vb, const double* sig2_beta_vb, const double* tau_vb,
                         const int* perm, int n_ind, double c, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > q) nthreads = q > 0 ? q : 1;
  primal_args* args = (primal_args*)malloc(sizeof(primal_args) * (size_t)nthreads);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  for (int t = 0; t < nthreads; ++t) {
    primal_args a = {n, p, q, (int)((long long)q * t / nthreads), (int)((long long)q * (t + 1) / nthreads),
                     n_ind, X, xnorm2, log_Phi, log_1_min_Phi, log_tau_vb, sig2_beta_vb, tau_vb,
                     R, gam_vb, m1_beta, mu_beta_vb, log_sig2_inv_vb, c, perm};
    args[t] = a;
  }
  if (nthreads == 1) {
    primal_range(&args[0]);
  } else {
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, primal_range, &args[t]);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  }
  free(args); free(th);
}

/* Blocked primal sweep with exact in-block Gauss-Seidel. */
void oracle_sweep_primal_blocked(int n, int p, int q, int B, const double* X, double* R, double* gam_vb,
                                 const double* log_Phi, const double* log_1_min_Phi,
                                 double log_sig2_inv_vb, const double* log_tau_vb, double* m1_beta,
                                 double* mu_beta_vb, const double* sig2_beta_vb, const double* tau_vb,
                                 const int* perm, int n_ind, double c) {
  double* G = (double*)malloc(sizeof(double) * (size_t)B * B);
  double* S = (double*)malloc(sizeof(double) * (size_t)B);
  double* D = (double*)malloc(sizeof(double) * (size_t)B);
  for (int b0 = 0; b0 < n_ind; b0 += B) {
    int nb = (n_ind - b0 < B) ? n_ind - b0 : B;
    for (int t = 0; t < nb; ++t)
      for (int u = 0; u <= t; ++u) {
        const double* xt = X + (size_t)perm[b0 + t] * n;
        const double* xu = X + (size_t)perm[b0 + u] * n;
        double g = 0;
        for (int i = 0; i < n; ++i) g += xt[i] * xu[i];
        G[t * B + u] = G[u * B + t] = g;
      }
    for (int k = 0; k < q; ++k) {
      double cst = -(log_tau_vb[k] + log_sig2_inv_vb + log(sig2_beta_vb[k])) / 2;
      double* r = R + (size_t)k * n;
      for (int t = 0; t < nb; ++t) {
        const double* x = X + (size_t)perm[b0 + t] * n;
        double dot = 0;
        for (int i = 0; i < n; ++i) dot += x[i] * r[i];
        S[t] = dot;
      }
      for (int t = 0; t < nb; ++t) {
        int j = perm[b0 + t];
        size_t jk = (size_t)j + (size_t)k * p;
        double s = S[t];
        for (int u = 0; u < t; ++u) s -= G[t * B + u] * D[u];
        double m1_old = m1_beta[jk];
        s += m1_old * G[t * B + t];
        double m1_new = pair_update(s, c, sig2_beta_vb[k], tau_vb[k],
                                    log_1_min_Phi[jk] - log_Phi[jk], cst, &mu_beta_vb[jk], &gam_vb[jk]);
        m1_beta[jk] = m1_new;
        D[t] = m1_new - m1_old;
      }
      for (int t = 0; t < nb; ++t) {
        const double* x = X + (size_t)perm[b0 + t] * n;
        double d = D[t];
        for (int i = 0; i < n; ++i) r[i] -= d * x[i];
      }
    }
  }
  free(G); free(S); free(D);
}

/* R = Y - X beta (n x q), beta p x q.  Used to initialise / re-derive the residual. */
void oracle_residual(int n, int p, int q, const double* X, const double* Y, const double* beta,
                     double* R) {
  for (int k = 0; k < q; ++k) {
    double* r = R + (size_t)k * n;
    memcpy(r, Y + (size_t)k * n, sizeof(double) * (size_t)n);
    for (int j = 0; j < p; ++j) {
      double b = beta[(size_t)j + (size_t)k * p];
      if (b == 0.0) continue;
      const double* x = X + (size_t)j * n;
      for (int i = 0; i < n; ++i) r[i] -= b * x[i];
    }
  }
}

/* Missing responses: coreDualMisLoop (reference src/coreLoop.cpp:91-138) in sample space.
 * mis n x q holds 1 where y_ik is observed, 0 where it is missing (R/atlasqtl_global_local_core.R:21);
 * the reference's per-trait Gram cp_X - cp_X_rm[[k]] is X' diag(mis_k) X (:25-32, src/coreLoop.cpp:120,132), so with a
 * residual kept at zero in the missing rows, r_k = mis_k o (y_k - X beta_k):
 *   cp_Y_X(k,j) - cp_betaX_X_jk = x_j' r_k + beta_jk * X_norm_sq(j,k),   X_norm_sq = crossprod(X^2, mis) (:23)
 *   r_k -= delta * (mis_k o x_j).
 * sig2_beta_vb is p x q here (update_sig2_beta_vb_, R/update_vb.R:47); cst has no log(sig2_beta) term (:108), which
 * moves into the exponent (:129). */
void oracle_sweep_primal_mis(int n, int p, int q, const double* X, const double* mis, const double* xnsq,
                             double* R, double* gam_vb, const double* log_Phi, const double* log_1_min_Phi,
                             double log_sig2_inv_vb, const double* log_tau_vb, double* m1_beta,
                             double* mu_beta_vb, const double* sig2_beta_vb, const double* tau_vb,
                             const int* perm, int n_ind, double c) {
  for (int k = 0; k < q; ++k) {
    double cst = -(log_tau_vb[k] + log_sig2_inv_vb) / 2;            /* :108 */
    double* r = R + (size_t)k * n;
    const double* m = mis + (size_t)k * n;
    for (int b = 0; b < n_ind; ++b) {
      int j = perm[b];
      size_t jk = (size_t)j + (size_t)k * p;
      const double* x = X + (size_t)j * n;
      double dot = 0;
      for (int i = 0; i < n; ++i) dot += x[i] * r[i];
      double m1_old = m1_beta[jk];
      double s = dot + m1_old * xnsq[jk];                            /* :120, :125 */
      double s2 = sig2_beta_vb[jk];
      double mu = c * s2 * tau_vb[k] * s;                            /* :125 */
      double gam = exp(-log_one_plus_exp(c * (log_1_min_Phi[jk] - log_Phi[jk] - mu * mu / (2 * s2) -
                                              log(s2) / 2 + cst)));  /* :127-129 */
      mu_beta_vb[jk] = mu;
      gam_vb[jk] = gam;
      double m1_new = gam * mu;                                      /* :131 */
      m1_beta[jk] = m1_new;
      double delta = m1_new - m1_old;
      for (int i = 0; i < n; ++i) r[i] -= delta * x[i] * m[i];       /* :132 */
    }
  }
}
