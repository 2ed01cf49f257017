"""Hyperparameter and initial-value objects: host-side mirror of R/set_hyper_init.R.

`set_hyper` / `set_init` keep the reference's argument names and checks
(R/set_hyper_init.R:98-140, :311-351); `auto_set_hyper_` / `auto_set_init_` restate the
defaults (:146-197, :356-418).  The draws use NumPy's generator: R's RNG stream cannot be
reproduced outside R, so bit-identical *default* inits are not promised -- parity runs pass
explicit `list_hyper` / `list_init` objects, exactly as the reference allows.
"""
import numpy as np
from scipy import optimize, special, stats


def _E_Phi_X(mu, s2):
    return stats.norm.cdf(mu / np.sqrt(1 + s2))  # R/utils.R:218-222


def _E_Phi_X_2(mu, s2):
    a = mu / np.sqrt(1 + s2)
    return stats.norm.cdf(a) - 2 * special.owens_t(a, 1 / np.sqrt(1 + 2 * s2))  # R/utils.R:224-229


def get_V_p_t(mu, s2, p):
    e1 = _E_Phi_X(mu, s2)
    return p * (p - 1) * _E_Phi_X_2(mu, s2) - p ** 2 * e1 ** 2 + p * e1  # R/utils.R:231-235


def get_mu(E_p_t, s2, p):
    return np.sqrt(1 + s2) * stats.norm.ppf(E_p_t / p)  # R/utils.R:236-240


def _zeroin(f, a, b, tol, maxit=1000):
    """Brent's zeroin as stats::uniroot runs it (R_zeroin2: linear / inverse quadratic interpolation with bisection
    safeguards, stop when the bracket is shorter than 4 eps |b| + tol), so that root = "uniroot" lands on the very
    iterate R returns."""
    eps = np.finfo(np.float64).eps
    fa, fb = f(a), f(b)
    if not (np.isfinite(fa) and np.isfinite(fb)) or fa * fb > 0:
        raise ValueError("f() values at end points not of opposite sign")
    if fa == 0.0:
        return a
    if fb == 0.0:
        return b
    c, fc = a, fa
    for _ in range(maxit + 1):
        prev_step = b - a
        if abs(fc) < abs(fb):
            a, b, c = b, c, b
            fa, fb, fc = fb, fc, fb
        tol_act = 2 * eps * abs(b) + tol / 2
        new_step = (c - b) / 2
        if abs(new_step) <= tol_act or fb == 0.0:
            return b
        if abs(prev_step) >= tol_act and abs(fa) > abs(fb):
            cb = c - b
            if a == c:
                t1 = fb / fa
                pp, qq = cb * t1, 1.0 - t1
            else:
                qq, t1, t2 = fa / fc, fb / fc, fb / fa
                pp = t2 * (cb * qq * (qq - t1) - (b - a) * (t1 - 1.0))
                qq = (qq - 1.0) * (t1 - 1.0) * (t2 - 1.0)
            if pp > 0:
                qq = -qq
            else:
                pp = -pp
            if pp < (0.75 * cb * qq - abs(tol_act * qq) / 2) and pp < abs(prev_step * qq / 2):
                new_step = pp / qq
        if abs(new_step) < tol_act:
            new_step = tol_act if new_step > 0 else -tol_act
        a, fa = b, fb
        b += new_step
        fb = f(b)
        if (fb > 0 and fc > 0) or (fb < 0 and fc < 0):
            c, fc = a, fa
    return b


def _solve_t02(p, p0, root="brentq"):
    """uniroot on [1e-6, 1e5] (R/set_hyper_init.R:163-175).  root = "brentq" (default) solves the equation to 1e-12;
    root = "uniroot" stops where R does (default tol = .Machine$double.eps^0.25 = 1.2e-4 on t02) and so reproduces
    the reference's t02 / n0 to rounding (tests/test_rlite.py) -- the two differ by up to that tolerance."""
    E_p_t, V_p_t = float(p0[0]), float(p0[1])
    f = lambda x: float(get_V_p_t(get_mu(E_p_t, x, p), x, p) - V_p_t)
    try:
        if root == "uniroot":
            return _zeroin(f, 1e-6, 1e5, np.finfo(np.float64).eps ** 0.25)
        if root != "brentq":
            raise TypeError(f"root must be 'brentq' or 'uniroot', not {root!r}")
        return optimize.brentq(f, 1e-6, 1e5, xtol=1e-12, rtol=1e-10)
    except ValueError:
        raise ValueError("No hyperparameter values matching the expectation and variance of the "
                         "number of active predictors per responses supplied in p0. Please change p0.")


def _check_positive(x, name):
    if np.any(np.asarray(x) <= 0):
        raise ValueError(f"{name} must be positive.")


def set_hyper(q, p, eta, kappa, n0, nu, rho, t02):
    """R/set_hyper_init.R:98-140."""
    q, p = int(q), int(p)
    n0 = np.full(q, float(n0)) if np.ndim(n0) == 0 else np.asarray(n0, float)
    eta = np.full(q, float(eta)) if np.ndim(eta) == 0 else np.asarray(eta, float)
    kappa = np.full(q, float(kappa)) if np.ndim(kappa) == 0 else np.asarray(kappa, float)
    for v, nm in ((n0, "n0"), (eta, "eta"), (kappa, "kappa")):
        if v.shape != (q,):
            raise ValueError(f"{nm} must have length 1 or q.")
    for v, nm in ((t02, "t02"), (nu, "nu"), (rho, "rho"), (eta, "eta"), (kappa, "kappa")):
        _check_positive(v, nm)
    return dict(q_hyper=q, p_hyper=p, A2_inv=1.0, eta=eta, kappa=kappa, m0=0.0, n0=n0, nu=float(nu),
                rho=float(rho), t02=float(t02), _class="hyper")


def auto_set_hyper_(Y, p, p0, root="brentq"):
    """R/set_hyper_init.R:146-197.  root: see `_solve_t02`."""
    q = Y.shape[1]
    eta = 1 / np.median(np.nanvar(Y, axis=0, ddof=1))  # apply(Y, 2, var, na.rm = TRUE)
    if not np.isfinite(eta):
        eta = 1e3
    t02 = _solve_t02(p, p0, root)
    n0 = get_mu(p0[0], t02, p)
    h = set_hyper(q, p, eta, 1.0, n0, 1e-2, 1.0, t02)
    h["_class"] = "out_hyper"
    return h


def set_init(q, p, gam_vb, mu_beta_vb, sig02_inv_vb, sig2_beta_vb, sig2_theta_vb, tau_vb, theta_vb, zeta_vb):
    """R/set_hyper_init.R:311-351."""
    q, p = int(q), int(p)
    gam_vb = np.asarray(gam_vb, float)
    mu_beta_vb = np.asarray(mu_beta_vb, float)
    if gam_vb.shape != (p, q) or mu_beta_vb.shape != (p, q):
        raise ValueError("gam_vb and mu_beta_vb must be p x q matrices.")
    if np.any(gam_vb < 0) or np.any(gam_vb > 1):
        raise ValueError("gam_vb must lie in [0, 1].")
    for v, nm, ln in ((sig2_beta_vb, "sig2_beta_vb", q), (sig2_theta_vb, "sig2_theta_vb", p), (tau_vb, "tau_vb", q)):
        if np.shape(v) != (ln,):
            raise ValueError(f"{nm} has the wrong length.")
        _check_positive(v, nm)
    _check_positive(sig02_inv_vb, "sig02_inv_vb")
    if np.shape(theta_vb) != (p,) or np.shape(zeta_vb) != (q,):
        raise ValueError("theta_vb / zeta_vb have the wrong length.")
    return dict(q_init=q, p_init=p, gam_vb=gam_vb, mu_beta_vb=mu_beta_vb, sig02_inv_vb=float(sig02_inv_vb),
                sig2_beta_vb=np.asarray(sig2_beta_vb, float), sig2_theta_vb=np.asarray(sig2_theta_vb, float),
                tau_vb=np.asarray(tau_vb, float), theta_vb=np.asarray(theta_vb, float),
                zeta_vb=np.asarray(zeta_vb, float), _class="init")


def auto_set_init_(Y, p, p0, shr_fac_inv, user_seed=None, root="brentq"):
    """R/set_hyper_init.R:356-418 (same distributions, NumPy generator)."""
    q = Y.shape[1]
    rng = np.random.default_rng(user_seed)
    t02 = _solve_t02(p, p0, root)
    n0 = get_mu(p0[0], t02, p)
    s02 = 1e-4
    gam_vb = stats.norm.cdf(rng.normal(n0, s02 + t02, size=(p, q)))
    mu_beta_vb = rng.normal(size=(p, q))
    sig2_inv_vb = 1e-2
    tau = 1 / np.median(np.nanvar(Y, axis=0, ddof=1))  # apply(Y, 2, var, na.rm = TRUE)
    if not np.isfinite(tau):
        tau = 1e3
    tau_vb = np.full(q, tau)
    sig2_beta_vb = 1 / rng.gamma(shape=2.0, scale=sig2_inv_vb * tau_vb)  # rate = 1/(sig2_inv*tau)
    sig02_inv_vb = rng.gamma(shape=max(p, q), scale=1.0)
    theta_vb = rng.normal(0.0, 1 / np.sqrt(sig02_inv_vb * shr_fac_inv), size=p)
    sig2_theta_vb = 1 / (q + rng.gamma(shape=sig02_inv_vb * shr_fac_inv, scale=1.0, size=p))
    zeta_vb = rng.normal(n0, np.sqrt(t02), size=q)
    out = set_init(q, p, gam_vb, mu_beta_vb, sig02_inv_vb, sig2_beta_vb, sig2_theta_vb, tau_vb, theta_vb, zeta_vb)
    out["_class"] = "out_init"
    return out
