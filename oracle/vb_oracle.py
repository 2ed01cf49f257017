"""NumPy/SciPy restatement of the reference's VB outer loop -- TEST INFRASTRUCTURE ONLY.

Follows, statement by statement, the global-local (horseshoe, df = 1) core of the reference:

    R/atlasqtl_global_local_core.R:8-433   atlasqtl_global_local_core_  (no-missing-value branch)
    R/atlasqtl_global_local_core.R:440-495 elbo_global_local_
    R/update_vb.R                          update_*_vb_ helpers
    R/elbo.R                               e_*_ ELBO terms
    R/utils.R:108-146, 172-191, 380-423    annealing ladder, inverse Mills ratio, Q_approx_vec

PARITY STATUS: pinned against the reference's own R sources as far as this image allows.  R itself (and Rcpp, gsl,
nmath) is absent, so the reference's unmodified R files are executed by the R evaluator of oracle/rlite (`.Call` bound
to the reference's own compiled src/coreLoop.cpp) and the outputs are committed as tests/golden/rlite_*.npz
(tests/golden/make_rlite_golden.py).  This restatement must reproduce them -- ELBO at every evaluation to 1e-12
relative, identical iteration counts, parameters to 1e-10, with and without annealing / thinning / missing responses,
and BASELINE config C1 -- in tests/test_rlite.py, which also re-derives the fixtures live where /root/reference exists.
What remains a stand-in: the evaluator (its R semantics are unit-tested) and SciPy for nmath / gsl (pinned against
mpmath in tests/test_special_functions.py).  The sweep it calls IS the reference's: `sweep="reference"` runs the
reference's own src/coreLoop.cpp (oracle/_ref).

Function mapping: pnorm(log.p=TRUE) -> scipy.special.log_ndtr;
digamma/lgamma -> scipy.special.digamma/gammaln; gsl::expint_E1 -> scipy.special.exp1;
gsl::gamma_inc(a, x) (unnormalised upper incomplete gamma, a > 0 on this path) ->
gamma(a) * gammaincc(a, x); all.equal(c, 1) -> |c - 1| < 1.5e-8.
"""
import numpy as np
from scipy import special as sp

from . import native

LOG_SQRT_2PI = np.log(np.sqrt(2 * np.pi))


# ----------------------------------------------------------------------------- R/utils.R
def get_annealing_ladder_(anneal):
    """R/utils.R:108-146."""
    k_m = 1.0 / anneal[1]
    m = int(anneal[2])
    seq_m1 = np.arange(m, 0, -1, dtype=np.float64)  # m:1
    if anneal[0] == 1:  # geometric
        delta_k = k_m ** (1.0 / (1 - m)) - 1
        return (1 + delta_k) ** (1 - seq_m1)
    if anneal[0] == 2:  # harmonic
        delta_k = (1 / k_m - 1) / (m - 1)
        return 1 / (1 + delta_k * (seq_m1 - 1))
    delta_k = (1 - k_m) / (m - 1)  # linear
    return k_m + delta_k * (np.arange(1, m + 1, dtype=np.float64) - 1)


def inv_mills_ratio_(y, U, log_1_pnorm_U, log_pnorm_U):
    """R/utils.R:172-191."""
    if y == 1:
        m = np.exp(-U ** 2 / 2 - LOG_SQRT_2PI - log_pnorm_U)
        m = np.where(m < -U, -U, m)
    else:
        m = -np.exp(-U ** 2 / 2 - LOG_SQRT_2PI - log_1_pnorm_U)
        m = np.where(m > -U, -U, m)
    return m


def Q_approx_vec(x, eps1=1e-30, eps2=1e-7):
    """R/utils.R:380-423 -- E1(x) exp(x) for x <= 1, modified Lentz continued fraction for x > 1
    with the reference's VECTOR-WIDE stopping rule max|Delta - 1| < eps2."""
    x = np.asarray(x, dtype=np.float64)
    out = np.full(x.shape, np.nan)
    lo = x <= 1
    if lo.any():
        out[lo] = sp.exp1(x[lo]) * np.exp(x[lo])
    up = ~lo
    if up.any():
        xu = x[up]
        f_p = np.full(xu.shape, eps1)
        C_p = np.full(xu.shape, eps1)
        D_p = np.zeros(xu.shape)
        Delta = np.full(xu.shape, 2 + eps2)
        j = 1
        f_c = f_p
        while np.max(np.abs(Delta - 1)) >= eps2:
            j += 1
            D_c = xu + 2 * j - 1 - ((j - 1) ** 2) * D_p
            C_c = xu + 2 * j - 1 - ((j - 1) ** 2) / C_p
            D_c = 1 / D_c
            Delta = C_c * D_c
            f_c = f_p * Delta
            f_p = f_c
            C_p = C_c
            D_p = D_c
        out[up] = 1 / (xu + 1 + f_c)
    return out


def gsl_gamma_inc(a, x):
    """gsl::gamma_inc(a, x): unnormalised upper incomplete gamma, a > 0 here."""
    return sp.gamma(a) * sp.gammaincc(a, x)


# ----------------------------------------------------------------------------- R/update_vb.R
def update_m2_beta_(gam_vb, mu_beta_vb, sig2_beta_vb):
    s2 = sig2_beta_vb[None, :] if np.ndim(sig2_beta_vb) == 1 else sig2_beta_vb  # p x q with missing responses
    return (mu_beta_vb ** 2 + s2) * gam_vb  # :19-31


def update_sig2_beta_vb_(n, sig2_inv_vb, tau_vb, c=1.0, X_norm_sq=None):
    if X_norm_sq is not None:
        return 1 / (c * (X_norm_sq + sig2_inv_vb) * tau_vb[None, :])  # :47 (missing responses: p x q)
    return 1 / (c * (n - 1 + sig2_inv_vb) * tau_vb)  # :33-50


def update_annealed_lam2_inv_vb_(L_vb, c, df):
    assert df == 1
    return gsl_gamma_inc(-c + 2, L_vb) / (gsl_gamma_inc(-c + 1, L_vb) * L_vb) - 1  # :70-75


def update_sig2_c0_vb_(d, s02, c=1.0):
    return 1 / (c * (d + (1 / s02)))  # :92


def update_zeta_vb_(colsums_Z, theta_vb, n0, sig2_zeta_vb, t02_inv, c=1.0):
    return c * sig2_zeta_vb * (colsums_Z + t02_inv * n0 - np.sum(theta_vb))  # :99-110


def update_nu_vb_(nu, sum_gam, c=1.0):
    return c * (nu + sum_gam / 2) - c + 1  # :116


def update_rho_vb_(rho, colsums_m2, tau_vb, c=1.0):
    return c * float(rho + np.dot(tau_vb, colsums_m2) / 2)  # :118


def update_log_sig2_inv_vb_(nu_vb, rho_vb):
    return sp.digamma(nu_vb) - np.log(rho_vb)  # :120


def update_eta_vb_(n, eta, colsums_gam, c=1.0):
    return c * (eta + n / 2 + colsums_gam / 2) - c + 1  # :127-134


def update_kappa_vb_dual_(n, Y_norm_sq, cp_Y_X, cp_X_Xbeta, kappa, beta_vb, m2_beta, sig2_inv_vb, c=1.0):
    """R/update_vb.R:136-146 exactly (dual quantities)."""
    diag_cp = np.sum(cp_X_Xbeta * beta_vb, axis=0)
    return c * (kappa + (Y_norm_sq - 2 * np.sum(beta_vb * cp_Y_X.T, axis=0)
                         + (n - 1 + sig2_inv_vb) * np.sum(m2_beta, axis=0)
                         + diag_cp - (n - 1) * np.sum(beta_vb ** 2, axis=0)) / 2)


def update_kappa_vb_primal_(n, resid_sq, kappa, colsums_beta2, colsums_m2, sig2_inv_vb, c=1.0):
    """Same quantity in sample space: |y_k|^2 - 2 b'X'y + b'X'Xb = |y_k - X b_k|^2."""
    return c * (kappa + (resid_sq + (n - 1 + sig2_inv_vb) * colsums_m2 - (n - 1) * colsums_beta2) / 2)


def update_kappa_vb_primal_mis_(resid_sq, kappa, X_norm_sq, beta_vb, m2_beta, sig2_inv_vb, c=1.0):
    """R/update_vb.R:149-154 (missing responses) in sample space: resid_sq is the squared norm of the masked residual,
    which equals Y_norm_sq - 2 colSums(beta * t(cp_Y_X)) + colSums(cp_X_Xbeta * beta) with the per-trait Gram."""
    return c * (kappa + (resid_sq + sig2_inv_vb * np.sum(m2_beta, axis=0) + np.sum(X_norm_sq * m2_beta, axis=0)
                         - np.sum(X_norm_sq * beta_vb ** 2, axis=0)) / 2)


def update_log_tau_vb_(eta_vb, kappa_vb):
    return sp.digamma(eta_vb) - np.log(kappa_vb)  # :159


def update_theta_vb_(rowsums_Z, m0, sig02_inv, sig2_theta_vb, zeta_vb, c=1.0):
    return c * sig2_theta_vb * (rowsums_Z + sig02_inv * m0 - np.sum(zeta_vb))  # :166-181


def update_Z_(gam_vb, mat_v_mu, log_1_pnorm, log_pnorm, c=1.0):
    """R/update_vb.R:217-234."""
    if not abs(c - 1) < 1.5e-8:
        sqrt_c = np.sqrt(c)
        log_pnorm = sp.log_ndtr(sqrt_c * mat_v_mu)
        log_1_pnorm = sp.log_ndtr(-sqrt_c * mat_v_mu)
    else:
        sqrt_c = 1.0
    imr0 = inv_mills_ratio_(0, sqrt_c * mat_v_mu, log_1_pnorm, log_pnorm)
    imr1 = inv_mills_ratio_(1, sqrt_c * mat_v_mu, log_1_pnorm, log_pnorm)
    return (gam_vb * (imr1 - imr0) + imr0) / sqrt_c + mat_v_mu


# ----------------------------------------------------------------------------- R/elbo.R
def e_beta_gamma_(gam_vb, log_1_pnorm, log_pnorm, log_sig2_inv_vb, log_tau_vb, m2_beta, sig2_beta_vb,
                  sig2_zeta_vb, sig2_theta_vb, sig2_inv_vb, tau_vb):
    """R/elbo.R:10-34."""
    eps = np.finfo(np.float64).eps ** 0.75
    arg = (log_sig2_inv_vb * gam_vb / 2 + gam_vb * log_tau_vb[None, :] / 2
           - m2_beta * tau_vb[None, :] * sig2_inv_vb / 2 + gam_vb * log_pnorm
           + (1 - gam_vb) * log_1_pnorm - sig2_zeta_vb / 2 - gam_vb * np.log(gam_vb + eps)
           - (1 - gam_vb) * np.log(1 - gam_vb + eps) - sig2_theta_vb[:, None] / 2)
    ls2 = (np.log(sig2_beta_vb) + 1)[None, :] if np.ndim(sig2_beta_vb) == 1 else np.log(sig2_beta_vb) + 1  # :28-33
    return float(np.sum(arg + 0.5 * gam_vb * ls2))


def e_sig2_inv_(nu, nu_vb, log_sig2_inv_vb, rho, rho_vb, sig2_inv_vb):
    return ((nu - nu_vb) * log_sig2_inv_vb - (rho - rho_vb) * sig2_inv_vb + nu * np.log(rho)
            - nu_vb * np.log(rho_vb) - sp.gammaln(nu) + sp.gammaln(nu_vb))  # :41-46


def e_sig2_inv_hs_(xi_inv_vb, nu_s0_vb, log_xi_inv_vb, log_sig02_inv_vb, rho_s0_vb, sig02_inv_vb):
    return (-0.5 * log_sig02_inv_vb - xi_inv_vb * sig02_inv_vb + log_xi_inv_vb / 2 - sp.gammaln(0.5)
            - (nu_s0_vb - 1) * log_sig02_inv_vb + rho_s0_vb * sig02_inv_vb
            - nu_s0_vb * np.log(rho_s0_vb) + sp.gammaln(nu_s0_vb))  # :49-56


def e_tau_(eta, eta_vb, kappa, kappa_vb, log_tau_vb, tau_vb):
    return float(np.sum((eta - eta_vb) * log_tau_vb - (kappa - kappa_vb) * tau_vb + eta * np.log(kappa)
                        - eta_vb * np.log(kappa_vb) - sp.gammaln(eta) + sp.gammaln(eta_vb)))  # :63-68


def e_theta_hs_(lam2_inv_vb, L_vb, log_sig02_inv_vb, m0, theta_vb, Q_app, sig02_inv_vb, sig2_theta_vb, df):
    assert df == 1
    return float(np.sum(log_sig02_inv_vb / 2 - sig02_inv_vb * lam2_inv_vb
                        * (theta_vb ** 2 + sig2_theta_vb - 2 * m0 * theta_vb + m0 ** 2) / 2
                        + (np.log(sig2_theta_vb) + 1) / 2 - np.log(np.pi) + L_vb * lam2_inv_vb
                        + np.log(Q_app)))  # :88-92


def e_y_(n, kappa, kappa_vb, log_tau_vb, colsums_m2, sig2_inv_vb, tau_vb):
    """n: scalar, or colSums(mis_pat) with missing responses (:140-142 -- the same expression)."""
    arg = -n / 2 * np.log(2 * np.pi) + n / 2 * log_tau_vb
    return float(np.sum(arg - tau_vb * (kappa_vb - colsums_m2 * sig2_inv_vb / 2 - kappa)))  # :135-146


def e_zeta_(zeta_vb, n0, sig2_zeta_vb, t02_inv, vec_sum_log_det_zeta):
    q = len(zeta_vb)
    return float((vec_sum_log_det_zeta - t02_inv * np.sum((zeta_vb - n0) ** 2)
                  - q * t02_inv * sig2_zeta_vb + q) / 2)  # :153-161


# ----------------------------------------------------------------------------- the core
class _SweepState:
    """Holds whatever the chosen form of the sweep needs between iterations."""

    def __init__(self, X, Y, beta_vb, form):
        self.form = form
        self.X = X
        self.n, self.p = X.shape
        if form in ("reference", "dual"):
            self.cp_X = np.asfortranarray(X.T @ X)  # R/atlasqtl_global_local_core.R:41
            self.cp_Y_X = np.asfortranarray(Y.T @ X)  # :42
            self.cp_X_Xbeta = np.asfortranarray(self.cp_X.T @ beta_vb)  # :115, update_vb.R:54-63
            self.Y_norm_sq = np.sum(Y ** 2, axis=0)  # :40
        else:
            self.xnorm2 = np.asfortranarray(np.sum(X ** 2, axis=0))
            self.R = native.residual(X, Y, beta_vb)
        self.mis = None

    def set_missing(self, mis_pat):
        """R/atlasqtl_global_local_core.R:19-33 in sample space: masked residual, X_norm_sq = crossprod(X^2, mis_pat)."""
        self.form = "primal_mis"
        self.mis = np.asfortranarray(mis_pat)
        self.X_norm_sq = np.asfortranarray((self.X ** 2).T @ self.mis)
        self.n_obs = self.mis.sum(axis=0)
        self.R = np.asfortranarray(self.R * self.mis)


def atlasqtl_global_local_core_(Y, X, shr_fac_inv, anneal, df, tol, maxit, list_hyper, list_init,
                                thinned_elbo_eval=True, debug=True, sweep="reference", block=8,
                                perm_fn=None, trace=None, nthreads=1, max_elbo_decrease=None):
    """Restatement of R/atlasqtl_global_local_core.R:8-433 (batch = "y", no missing values).

    Y (n x q, centred), X (n x p, standardised): Fortran float64.  list_hyper / list_init: dicts
    with the fields of R/set_hyper_init.R:133-136 / :344-348.  `sweep` selects the form of step 10:
    "reference" (the reference's own coreLoop.cpp), "dual", "primal", "blocked" (cavi_oracle.c).
    `perm_fn(it, p)` returns shuffled_ind for iteration `it` (default: identity, as the reference,
    R/atlasqtl_global_local_core.R:162).  `trace`, if a list, receives one dict per iteration.
    """
    Y = np.asfortranarray(Y, dtype=np.float64)
    X = np.asfortranarray(X, dtype=np.float64)
    n, q = Y.shape
    p = X.shape[1]
    mis_pat = None
    if np.isnan(Y).any():  # :19-33
        mis_pat = np.where(np.isnan(Y), 0.0, 1.0)
        Y = np.asfortranarray(np.where(np.isnan(Y), 0.0, Y))
        sweep = "primal"
    h = list_hyper
    eta, kappa, n0, nu, rho, t02 = (np.asarray(h["eta"], float), np.asarray(h["kappa"], float),
                                    np.asarray(h["n0"], float), float(h["nu"]), float(h["rho"]),
                                    float(h["t02"]))
    m0, A2_inv = float(h.get("m0", 0.0)), float(h.get("A2_inv", 1.0))

    gam_vb = np.array(list_init["gam_vb"], dtype=np.float64, order="F")
    mu_beta_vb = np.array(list_init["mu_beta_vb"], dtype=np.float64, order="F")
    sig02_inv_vb = float(list_init["sig02_inv_vb"])
    sig2_beta_vb = np.array(list_init["sig2_beta_vb"], dtype=np.float64)
    sig2_theta_vb = np.array(list_init["sig2_theta_vb"], dtype=np.float64)
    tau_vb = np.array(list_init["tau_vb"], dtype=np.float64)
    theta_vb = np.array(list_init["theta_vb"], dtype=np.float64)
    zeta_vb = np.array(list_init["zeta_vb"], dtype=np.float64)

    theta_plus_zeta_vb = np.asfortranarray(theta_vb[:, None] + zeta_vb[None, :])  # :61
    log_Phi = np.asfortranarray(sp.log_ndtr(theta_plus_zeta_vb))  # :62
    log_1_min_Phi = np.asfortranarray(sp.log_ndtr(-theta_plus_zeta_vb))  # :63

    anneal_scale = True  # :71
    if anneal is None:
        annealing, c, c_s, it_init = False, 1.0, 1.0, 1
        ladder = None
    else:
        annealing = True
        ladder = get_annealing_ladder_(anneal)
        c = float(ladder[0])
        c_s = c if anneal_scale else 1.0
        it_init = int(anneal[2])

    eps = np.finfo(np.float64).eps ** 0.5  # :85
    if thinned_elbo_eval:
        times_conv_sched = np.array([1, 5, 10, 50], float)
        batch_conv_sched = [1, 10, 25, 50]
    else:
        times_conv_sched = np.array([1.0])
        batch_conv_sched = [1]
    ind_batch_conv = len(batch_conv_sched) + 1
    batch_conv = 1

    t02_inv = 1 / t02
    sig2_zeta_vb = update_sig2_c0_vb_(p, t02, c=c)  # :105
    vec_sum_log_det_zeta = -q * (np.log(t02) + np.log(p + t02_inv))  # :107

    beta_vb = np.asfortranarray(gam_vb * mu_beta_vb)  # :112
    m2_beta = update_m2_beta_(gam_vb, mu_beta_vb, sig2_beta_vb)  # :113
    st = _SweepState(X, Y, beta_vb, sweep)
    if mis_pat is not None:
        st.set_missing(mis_pat)
    n_eff = n if mis_pat is None else st.n_obs  # update_eta_vb_ / e_y_ use colSums(mis_pat) (R/update_vb.R:131, R/elbo.R:141)
    nu_xi_inv_vb = 1.0  # :119

    converged = False
    lb_new = -np.inf
    lb_old = -np.inf
    it = 0
    Q_app = None
    sample_q = np.arange(q, dtype=np.int32)

    while (not converged) and (it < maxit):
        lb_old = lb_new
        it += 1

        colsums_m2 = np.sum(m2_beta, axis=0)
        nu_vb = update_nu_vb_(nu, np.sum(gam_vb), c=c)  # :134
        rho_vb = update_rho_vb_(rho, colsums_m2, tau_vb, c=c)  # :135 (OLD tau)
        sig2_inv_vb = nu_vb / rho_vb  # :137

        eta_vb = update_eta_vb_(n_eff, eta, np.sum(gam_vb, axis=0), c=c)  # :141
        if st.form == "primal_mis":
            kappa_vb = update_kappa_vb_primal_mis_(np.sum(st.R ** 2, axis=0), kappa, st.X_norm_sq, beta_vb, m2_beta,
                                                   sig2_inv_vb, c=c)
        elif st.form in ("reference", "dual"):
            kappa_vb = update_kappa_vb_dual_(n, st.Y_norm_sq, st.cp_Y_X, st.cp_X_Xbeta, kappa, beta_vb,
                                             m2_beta, sig2_inv_vb, c=c)  # :142
        else:
            kappa_vb = update_kappa_vb_primal_(n, np.sum(st.R ** 2, axis=0), kappa,
                                               np.sum(beta_vb ** 2, axis=0), colsums_m2, sig2_inv_vb, c=c)
        tau_vb = eta_vb / kappa_vb  # :145
        sig2_beta_vb = update_sig2_beta_vb_(n, sig2_inv_vb, tau_vb, c=c,
                                            X_norm_sq=st.X_norm_sq if st.form == "primal_mis" else None)  # :147
        log_tau_vb = update_log_tau_vb_(eta_vb, kappa_vb)  # :149
        log_sig2_inv_vb = float(update_log_sig2_inv_vb_(nu_vb, rho_vb))  # :150

        shuffled_ind = (np.arange(p, dtype=np.int32) if perm_fn is None
                        else np.ascontiguousarray(perm_fn(it, p), dtype=np.int32))  # :162

        # ---- step 10: the sweep (:167-170) ----
        if st.form in ("reference", "dual"):
            native.core_dual_loop(st.cp_X, st.cp_Y_X, gam_vb, log_Phi, log_1_min_Phi, log_sig2_inv_vb,
                                  log_tau_vb, beta_vb, st.cp_X_Xbeta, mu_beta_vb, sig2_beta_vb, tau_vb,
                                  shuffled_ind, sample_q, c=c,
                                  impl="reference" if st.form == "reference" else "oracle")
        elif st.form == "primal_mis":  # coreDualMisLoop (:172-175)
            native.sweep_primal_mis(st.X, st.mis, st.X_norm_sq, st.R, gam_vb, log_Phi, log_1_min_Phi, log_sig2_inv_vb,
                                    log_tau_vb, beta_vb, mu_beta_vb, np.asfortranarray(sig2_beta_vb), tau_vb,
                                    shuffled_ind, c=c)
        elif st.form == "primal":
            native.sweep_primal(st.X, st.xnorm2, st.R, gam_vb, log_Phi, log_1_min_Phi, log_sig2_inv_vb,
                                log_tau_vb, beta_vb, mu_beta_vb, sig2_beta_vb, tau_vb, shuffled_ind,
                                c=c, nthreads=nthreads)
        elif st.form == "blocked":
            native.sweep_primal_blocked(st.X, st.R, gam_vb, log_Phi, log_1_min_Phi, log_sig2_inv_vb,
                                        log_tau_vb, beta_vb, mu_beta_vb, sig2_beta_vb, tau_vb,
                                        shuffled_ind, c=c, B=block)
        else:
            raise ValueError(sweep)

        m2_beta = update_m2_beta_(gam_vb, mu_beta_vb, sig2_beta_vb)  # :235
        Z = update_Z_(gam_vb, theta_plus_zeta_vb, log_1_min_Phi, log_Phi, c=c)  # :237
        rowsums_Z = np.sum(Z, axis=1)
        colsums_Z = np.sum(Z, axis=0)

        L_vb = c_s * sig02_inv_vb * shr_fac_inv * (theta_vb ** 2 + sig2_theta_vb - 2 * theta_vb * m0
                                                   + m0 ** 2) / 2 / df  # :241
        rho_xi_inv_vb = c_s * (A2_inv + sig02_inv_vb)  # :242
        if annealing and anneal_scale:
            lam2_inv_vb = update_annealed_lam2_inv_vb_(L_vb, c_s, df)  # :246
        else:
            Q_app = Q_approx_vec(L_vb)  # :250
            lam2_inv_vb = 1 / (Q_app * L_vb) - 1  # :254
        xi_inv_vb = nu_xi_inv_vb / rho_xi_inv_vb  # :276
        sig2_theta_vb = update_sig2_c0_vb_(q, 1 / (sig02_inv_vb * lam2_inv_vb * shr_fac_inv), c=c)  # :278
        theta_vb = update_theta_vb_(rowsums_Z, m0, sig02_inv_vb * lam2_inv_vb * shr_fac_inv, sig2_theta_vb,
                                    zeta_vb, c=c)  # :280
        nu_s0_vb = update_nu_vb_(0.5, p, c=c_s)  # :283
        rho_s0_vb = c_s * (xi_inv_vb + np.sum(lam2_inv_vb * shr_fac_inv * (
            theta_vb ** 2 + sig2_theta_vb - 2 * theta_vb * m0 + m0 ** 2)) / 2)  # :285
        sig02_inv_vb = float(nu_s0_vb / rho_s0_vb)  # :288
        zeta_vb = update_zeta_vb_(colsums_Z, theta_vb, n0, sig2_zeta_vb, t02_inv, c=c)  # :290

        theta_plus_zeta_vb = np.asfortranarray(theta_vb[:, None] + zeta_vb[None, :])  # :293
        log_Phi = np.asfortranarray(sp.log_ndtr(theta_plus_zeta_vb))  # :294
        log_1_min_Phi = np.asfortranarray(sp.log_ndtr(-theta_plus_zeta_vb))  # :295

        rec = dict(it=it, c=c, annealing=annealing, lb=None, sig2_inv_vb=sig2_inv_vb,
                   sig02_inv_vb=sig02_inv_vb, sum_gam=float(np.sum(gam_vb)))
        if annealing:  # :318-336
            sig2_zeta_vb = c * sig2_zeta_vb
            c = float(ladder[it]) if it < len(ladder) else 1.0  # ladder[it + 1], R is 1-based
            c_s = c if anneal_scale else 1.0
            sig2_zeta_vb = sig2_zeta_vb / c
            if abs(c - 1) < 1.5e-8:
                annealing = False
        else:
            if it <= it_init + 1 or it % batch_conv == 0 or it % batch_conv == 1:  # :342
                lb_new = elbo_global_local_(n, p, A2_inv, df, eta, gam_vb, kappa, L_vb, lam2_inv_vb,
                                            log_1_min_Phi, log_Phi, m0, m2_beta, n0, nu, nu_s0_vb,
                                            nu_xi_inv_vb, Q_app, rho, rho_s0_vb, rho_xi_inv_vb, shr_fac_inv,
                                            sig02_inv_vb, sig2_beta_vb, sig2_inv_vb, sig2_theta_vb,
                                            sig2_zeta_vb, t02_inv, tau_vb, theta_vb, vec_sum_log_det_zeta,
                                            xi_inv_vb, zeta_vb, st, beta_vb)
                rec["lb"] = lb_new
                if debug and lb_new + eps < lb_old:  # :359-360
                    if max_elbo_decrease is None or lb_old - lb_new > max_elbo_decrease:
                        raise RuntimeError("ELBO not increasing monotonically. Exit. "
                                           f"(it={it}, lb_old={lb_old!r}, lb_new={lb_new!r})")
                diff_lb = abs(lb_new - lb_old)
                sum_exceed = int(np.sum(diff_lb > (times_conv_sched * tol)))  # :364
                if sum_exceed == 0:
                    converged = True
                elif ind_batch_conv > sum_exceed:
                    ind_batch_conv = sum_exceed
                    batch_conv = batch_conv_sched[ind_batch_conv - 1]
        if trace is not None:
            rec["max_gam"] = float(gam_vb.max())
            trace.append(rec)

    lb_opt = lb_new
    diff_lb = abs(lb_opt - lb_old)
    return dict(beta_vb=beta_vb, gam_vb=gam_vb, mu_beta_vb=mu_beta_vb, theta_vb=theta_vb, zeta_vb=zeta_vb,
                n=n, p=p, q=q, anneal=anneal, converged=converged, it=it, maxit=maxit, tol=tol,
                lb_opt=lb_opt, diff_lb=diff_lb, tau_vb=tau_vb, sig2_beta_vb=sig2_beta_vb,
                sig2_theta_vb=sig2_theta_vb, sig02_inv_vb=sig02_inv_vb, lam2_inv_vb=lam2_inv_vb)


def elbo_global_local_(n, p, A2_inv, df, eta, gam_vb, kappa, L_vb, lam2_inv_vb, log_1_min_Phi, log_Phi, m0,
                       m2_beta, n0, nu, nu_s0_vb, nu_xi_inv_vb, Q_app, rho, rho_s0_vb, rho_xi_inv_vb,
                       shr_fac_inv, sig02_inv_vb, sig2_beta_vb, sig2_inv_vb, sig2_theta_vb, sig2_zeta_vb,
                       t02_inv, tau_vb, theta_vb, vec_sum_log_det_zeta, xi_inv_vb, zeta_vb, st, beta_vb):
    """R/atlasqtl_global_local_core.R:440-495 (c = 1 re-derivations :456-467)."""
    colsums_m2 = np.sum(m2_beta, axis=0)
    if st.form == "primal_mis":
        n = st.n_obs
    eta_vb = update_eta_vb_(n, eta, np.sum(gam_vb, axis=0))
    if st.form == "primal_mis":
        kappa_vb = update_kappa_vb_primal_mis_(np.sum(st.R ** 2, axis=0), kappa, st.X_norm_sq, beta_vb, m2_beta,
                                               sig2_inv_vb)
    elif st.form in ("reference", "dual"):
        kappa_vb = update_kappa_vb_dual_(n, st.Y_norm_sq, st.cp_Y_X, st.cp_X_Xbeta, kappa, beta_vb, m2_beta,
                                         sig2_inv_vb)
    else:
        kappa_vb = update_kappa_vb_primal_(n, np.sum(st.R ** 2, axis=0), kappa,
                                           np.sum(beta_vb ** 2, axis=0), colsums_m2, sig2_inv_vb)
    nu_vb = update_nu_vb_(nu, np.sum(gam_vb))
    rho_vb = update_rho_vb_(rho, colsums_m2, tau_vb)
    log_tau_vb = update_log_tau_vb_(eta_vb, kappa_vb)
    log_sig2_inv_vb = update_log_sig2_inv_vb_(nu_vb, rho_vb)
    log_sig02_inv_vb = update_log_sig2_inv_vb_(nu_s0_vb, rho_s0_vb)
    log_xi_inv_vb = update_log_sig2_inv_vb_(nu_xi_inv_vb, rho_xi_inv_vb)

    A = e_y_(n, kappa, kappa_vb, log_tau_vb, colsums_m2, sig2_inv_vb, tau_vb)
    B = e_beta_gamma_(gam_vb, log_1_min_Phi, log_Phi, log_sig2_inv_vb, log_tau_vb, m2_beta, sig2_beta_vb,
                      sig2_zeta_vb, sig2_theta_vb, sig2_inv_vb, tau_vb)
    C = e_theta_hs_(lam2_inv_vb, L_vb, log_sig02_inv_vb + np.log(shr_fac_inv), m0, theta_vb, Q_app,
                    sig02_inv_vb * shr_fac_inv, sig2_theta_vb, df)
    D = e_zeta_(zeta_vb, n0, sig2_zeta_vb, t02_inv, vec_sum_log_det_zeta)
    E = e_tau_(eta, eta_vb, kappa, kappa_vb, log_tau_vb, tau_vb)
    F = e_sig2_inv_hs_(xi_inv_vb, nu_s0_vb, log_xi_inv_vb, log_sig02_inv_vb, rho_s0_vb, sig02_inv_vb)
    G = e_sig2_inv_(0.5, nu_xi_inv_vb, log_xi_inv_vb, A2_inv, rho_xi_inv_vb, xi_inv_vb)
    H = e_sig2_inv_(nu, nu_vb, log_sig2_inv_vb, rho, rho_vb, sig2_inv_vb)
    return float(A + B + C + D + E + F + G + H)


# ----------------------------------------------------------------------------- post-processing
def assign_bFDR(mat_ppi):
    """R/summarise_output.R:207-223 (column-major as.vector; stable descending order)."""
    vec = np.asarray(mat_ppi, dtype=np.float64).flatten(order="F")
    ind = np.argsort(-vec, kind="stable")
    fdr_ord = np.cumsum(1 - vec[ind]) / np.arange(1, len(vec) + 1)
    out = np.empty_like(vec)
    out[ind] = fdr_ord
    return out.reshape(mat_ppi.shape, order="F")
