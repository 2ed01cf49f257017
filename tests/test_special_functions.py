"""CPU: the SciPy stand-ins the restated outer loop uses for R nmath / GSL (oracle/vb_oracle.py, atlasqtl_b200/core.py;
SURVEY.md section 8c) pinned against independent high-precision evaluations (mpmath, 40 digits) over the argument ranges
the loop visits.  R itself is not in the image, so this is the strongest available check that
  pnorm(x, log.p = TRUE)      -> scipy.special.log_ndtr          (R/atlasqtl_global_local_core.R:62-63, :294-295)
  gsl::expint_E1(x) * exp(x)  -> scipy.special.exp1 * exp         (R/utils.R:350, :387; x <= 1 branch of Q_approx_vec)
  gsl::gamma_inc(a, x)        -> gamma(a) * gammaincc(a, x)       (R/update_vb.R:74; a = 2 - c, 1 - c with c in (0, 1))
  digamma / lgamma            -> scipy.special.digamma / gammaln  (R/update_vb.R:120, :159; R/elbo.R)
agree with the mathematical functions to the accuracy the parity bars need (ELBO relative 1e-10, gam_vb 1e-8)."""
import mpmath as mp
import numpy as np
from scipy import special as sp

mp.mp.dps = 40


def _rel(a, b):
    d = abs(mp.mpf(float(a)) - b)
    # (results below 1e-30 in magnitude -- log Phi(x) for x > 11 -- only ever enter sums of O(1) terms: absolute agreement)
    return 0.0 if d < mp.mpf("1e-45") else float(d / abs(b))


def test_log_ndtr_both_tails():
    # theta + zeta ranges from about -30 (null pairs after shrinkage) to +10 (hotspot pairs)
    with mp.workdps(400):   # log Phi(38) = -3e-316 needs more than 316 digits of Phi
        _log_ndtr_both_tails()


def _log_ndtr_both_tails():
    for x in np.concatenate([np.linspace(-38, -5, 67), np.linspace(-5, 5, 41), np.linspace(5, 12, 15)]):
        want = mp.log(mp.ncdf(mp.mpf(float(x))))
        assert _rel(sp.log_ndtr(x), want) < 5e-14, x   # (SciPy: a few 1e-14 relative in the far tails)
        want_u = mp.log(mp.ncdf(-mp.mpf(float(x))))
        assert _rel(sp.log_ndtr(-x), want_u) < 5e-14, x


def test_exp1_times_exp_small_arguments():
    for x in np.concatenate([10.0 ** np.linspace(-300, -1, 60), np.linspace(0.1, 1.0, 19)]):
        want = mp.e1(mp.mpf(float(x))) * mp.exp(mp.mpf(float(x)))
        assert _rel(sp.exp1(x) * np.exp(x), want) < 1e-14, x


def test_lentz_branch_accuracy_and_vector_coupling():
    """x > 1: the reference's modified Lentz loop stops on the VECTOR-WIDE criterion max|Delta - 1| < 1e-7
    (R/utils.R:402), so a value depends on its companions: truncation error up to ~1e-7 relative by design.  Both
    restatements (oracle and product) must carry exactly that behaviour."""
    from atlasqtl_b200 import core
    from oracle import vb_oracle
    x = np.array([1.5, 3.0, 10.0, 200.0, 5e4])
    for fn in (core.Q_approx_vec, vb_oracle.Q_approx_vec):
        got = fn(x)
        for xi, gi in zip(x, got):
            want = mp.e1(mp.mpf(float(xi))) * mp.exp(mp.mpf(float(xi)))
            assert _rel(gi, want) < 1e-6, xi
        # an argument alone stops after fewer steps than next to a slower-converging companion
        alone, paired = fn(np.array([4.0]))[0], fn(np.array([1.01, 4.0]))[1]
        assert alone != paired and abs(alone - paired) / paired < 1e-6
    np.testing.assert_array_equal(core.Q_approx_vec(x), vb_oracle.Q_approx_vec(x))


def test_upper_incomplete_gamma_on_the_annealing_ladder():
    from atlasqtl_b200 import core
    ladder = core.get_annealing_ladder_((1, 2, 10))   # c = 2^(-9/9) .. 1
    for c in ladder[:-1]:
        for a in (2 - c, 1 - c):
            for x in (1e-8, 1e-3, 0.3, 1.0, 7.0, 40.0, 300.0):
                want = mp.gammainc(mp.mpf(float(a)), mp.mpf(float(x)), mp.inf)
                got = sp.gamma(a) * sp.gammaincc(a, x)
                assert _rel(got, want) < 2e-13, (c, a, x)
    # the ratio the loop forms (R/update_vb.R:74) for a spread of L_vb
    L = np.array([1e-6, 1e-2, 0.5, 3.0, 50.0])
    for c in (0.5, 2 ** (-4 / 9)):
        got = core.update_annealed_lam2_inv_vb_(L, c, 1)
        for Li, gi in zip(L, got):
            Lm = mp.mpf(float(Li))
            want = mp.gammainc(2 - mp.mpf(float(c)), Lm, mp.inf) / (mp.gammainc(1 - mp.mpf(float(c)), Lm, mp.inf) * Lm) - 1
            assert _rel(gi, want) < 1e-11, (c, Li)


def test_digamma_gammaln():
    for x in (1e-2, 0.5, 1.0, 2.5, 100.5, 1e4 + 0.01, 5e6):
        assert _rel(sp.digamma(x), mp.digamma(mp.mpf(float(x)))) < 1e-14, x
        if abs(x - 1.0) > 1e-9:   # lgamma(1) = 0: compare absolutely there
            assert _rel(sp.gammaln(x), mp.loggamma(mp.mpf(float(x)))) < 1e-14, x
    assert abs(sp.gammaln(1.0)) < 1e-16


def test_logistic_forms_agree():
    """exp(-log1pexp(x)) as the reference writes it (src/coreLoop.cpp:28-33, :75-77) == 1 / (1 + exp(x))."""
    for x in (-745.0, -700.0, -40.0, -1.0, 0.0, 1e-300, 3.0, 40.0, 700.0, 745.0):
        m = max(x, 0.0)
        ref = np.exp(-(np.log(np.exp(x - m) + np.exp(-m)) + m))
        want = 1 / (1 + mp.exp(mp.mpf(float(x))))
        if want > mp.mpf("1e-300"):
            assert _rel(ref, want) < 1e-13, x   # (the log1pexp detour costs a few ulp for large |x|)
        else:
            assert ref < 1e-300
