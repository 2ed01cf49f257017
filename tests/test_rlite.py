"""CPU: the reference's OWN R code as the pin of the outer loop.

tests/golden/rlite_*.npz hold outputs of the reference's unmodified R sources (R/atlasqtl_global_local_core.R,
update_vb.R, elbo.R, utils.R, prepare_atlasqtl.R, summarise_output.R, set_hyper_init.R, atlasqtl.R) executed by the R
evaluator of oracle/rlite, with `.Call` bound to the reference's own src/coreLoop.cpp (tests/golden/make_rlite_golden.py).
Here:
  * the evaluator itself is checked on R semantics with known answers (recycling, indexing, replacement, scoping ...);
  * the NumPy restatement oracle/vb_oracle.py -- the checker of every GPU full-run test -- must reproduce those outputs
    (ELBO at every evaluation to 1e-12 relative, same iteration count, parameters to 1e-10);
  * the product's host loop (atlasqtl_b200.core, oracle-backed test double in place of the CUDA context), the
    pre-processing mirror, `atlasqtl()`, the hyper-parameter defaults and assign_bFDR are compared with them too;
  * where /root/reference exists (this container, not the GPU box) the fixtures are re-derived live.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)

from rlite_cases import CORE_FILES, CORE_IDS, hyper_init_of, load_case  # noqa: E402


# --------------------------------------------------------------------------------------------- the evaluator itself
R_SEMANTICS = [
    ("x <- c(1,2,3); x * 2 + 1", [3, 5, 7]),
    ("-2:2", [-2, -1, 0, 1, 2]),
    ("-(2:4)^2", [-4, -9, -16]),
    ("-2^2", [-4]),
    ("c(5 %% 3, -5 %/% 2, 7.5 %% 2, -5 %% 3)", [2, -3, 1.5, 1]),
    ("m <- matrix(1:6, nrow = 2); m[2, ]", [2, 4, 6]),
    ("m <- matrix(1:6, nrow = 2); as.vector(m * c(10, 100))", [10, 200, 30, 400, 50, 600]),        # column-major recycling
    ("m <- matrix(1:6, nrow = 2); as.vector(sweep(m, 2, c(1, 2, 3), `+`))", [2, 3, 5, 6, 8, 9]),
    ("m <- matrix(1:6, nrow = 2); as.vector(sweep(m, 1, c(10, 20), `*`))", [10, 40, 30, 80, 50, 120]),
    ("m <- matrix(1:6, nrow = 2); c(colSums(m), rowSums(m), sum(m))", [3, 7, 11, 9, 12, 21]),
    ("m <- matrix(1:6, nrow = 2); as.vector(t(m) %*% m)", [5, 11, 17, 11, 25, 39, 17, 39, 61]),
    ("as.vector(crossprod(matrix(1:6, nrow = 2), matrix(1:4, nrow = 2)))", [5, 11, 17, 11, 25, 39]),
    ("as.vector(tcrossprod(c(1, 2), rep(1, 3)))", [1, 2, 1, 2, 1, 2]),
    ("f <- function(a, b = a * 2, ...) a + b; c(f(1), f(1, 5), f(b = 1, a = 10))", [3, 6, 11]),     # lazy defaults
    ("x <- 1:10; x[x > 5] <- 0; x", [1, 2, 3, 4, 5, 0, 0, 0, 0, 0]),
    ("x <- 1:10; x[-1][c(TRUE, FALSE)]", [2, 4, 6, 8, 10]),
    ("x <- c(5, 3, 8); x[order(x, decreasing = TRUE)]", [8, 5, 3]),
    ("order(c(3, 1, 2, 3), decreasing = TRUE)", [1, 4, 3, 2]),                                      # ties: stable
    ("cumsum(c(1, 2, 3)) / 1:3", [1, 1.5, 2]),
    ("sapply(1:3, function(i) i^2)", [1, 4, 9]),
    ("as.vector(sapply(1:3, function(i) c(i, i^2)))", [1, 1, 2, 4, 3, 9]),
    ("unlist(lapply(1:3, function(i) i * 2))", [2, 4, 6]),
    ("if (FALSE) 1 else if (TRUE) 2 else 3", [2]),
    ("k <- 0; while (k < 5) { k <- k + 1; if (k == 3) break }; k", [3]),
    ("s <- 0; for (i in 1:10) { if (i %% 2 == 0) next; s <- s + i }; s", [25]),
    ("a <- b <- d <- NULL; c(is.null(b), is.null(1))", [1, 0]),
    ("c(isTRUE(all.equal(1, 1 + 1e-10)), isTRUE(all.equal(0.5, 1)), isTRUE(all.equal(1 - 1e-7, 1)))", [1, 0, 0]),
    ("with(list(u = 2, v = 3), { w <- u * v; w + 1 })", [7]),
    ("c <- 0.5; c(1, c)", [1, 0.5]),                                       # a variable named c does not hide c()
    ("sweep <- TRUE; as.vector(sweep(matrix(1, 1, 2), 2, c(1, 2), `+`))", [2, 3]),
    ("z <- c(1, 2, 3); z[5] <- 9; z", [1, 2, 3, np.nan, 9]),
    ("x <- rep(NA, 3); x[c(TRUE, FALSE, TRUE)] <- c(5, 6); x", [5, np.nan, 6]),
    ("(3 - 1)^{2}", [4]),
    ("c(seq(1, 7, by = 2), 3:1 - 1, seq_along(c(5, 6)))", [1, 3, 5, 7, 2, 1, 0, 1, 2]),
    ("as.numeric(crossprod(c(1, 2, 3), c(4, 5, 6)) / 2)", [16]),
    ("m <- c(1, -2, 3); m[m < 0] <- 0; m", [1, 0, 3]),
    ("Y <- matrix(c(1, NA, 3, 4), 2); Y[is.na(Y)] <- 0; as.vector(Y)", [1, 0, 3, 4]),
    ("x <- matrix(1:4, 2); x[x > 2]", [3, 4]),
    ("duplicated(matrix(c(1, 2, 1, 2, 3, 4), nrow = 2), MARGIN = 2)", [0, 1, 0]),
    ("duplicated(matrix(c(1, 2, 1, 2, 3, 4), nrow = 2), MARGIN = 2, fromLast = TRUE)", [1, 0, 0]),
    ("as.vector(scale(matrix(c(1, 2, 3, 4, 5, 9), ncol = 2)))", [-1, 0, 1, -2 / 7 ** 0.5, -1 / 7 ** 0.5, 3 / 7 ** 0.5]),
    ("as.vector(scale(matrix(c(1, NA, 3, 4), 2), center = TRUE, scale = FALSE))", [0, np.nan, -0.5, 0.5]),
    ("l <- list(a = 1, b = 2); l$c <- 3; l[['b']] <- 5; unlist(l)", [1, 5, 3]),
    ("ll <- lapply(1:2, function(k) matrix(k, 2, 2)); ll[[2]][1, ]", [2, 2]),
    ("x <- list(a = 1); class(x) <- 'hyper'; c(inherits(x, c('hyper', 'out_hyper')), inherits(x, 'init'))", [1, 0]),
    ("c(median(c(5, 1, 3, 2)), var(c(1, 2, 3, 4)), mean(c(1, 2, 4)))", [2.5, 5 / 3, 7 / 3]),
    ("m <- matrix(c(1, 2, 3, 4), 2); m[m > 1 & m < 4]", [2, 3]),
    ("h <- function(x) { if (x > 0) return(1); -1 }; c(h(1), h(-1))", [1, -1]),
    ("g <- function(x, y) { if (missing(y)) 0 else y }; c(g(1), g(1, 2))", [0, 2]),
    ("ifelse(c(1, 5, 3) > 2, 1, 0)", [0, 1, 1]),
    ("m <- matrix(0, 2, 2); rownames(m) <- c('r1', 'r2'); colnames(m) <- c('c1', 'c2'); m['r2', 'c1'] <- 5; as.vector(m)",
     [0, 5, 0, 0]),
    ("x <- c(a = 1, b = 2); names(x)[2] <- 'z'; x['z']", [2]),
    ("apply(matrix(1:6, 2), 2, function(v) sum(v))", [3, 7, 11]),
    ("which(c(FALSE, TRUE, TRUE))", [2, 3]),
    ("c(1 %in% c(0, 1), 3 %in% 0:2)", [1, 0]),
    ("c(pnorm(1.96), pnorm(-40, log.p = TRUE), pnorm(2, lower.tail = FALSE, log.p = TRUE))",
     [0.9750021048517795, -804.6084420137538, -3.7831843336820317]),
    ("c(.Machine$double.eps^0.5, digamma(1), lgamma(0.5), lfactorial(4))",
     [1.4901161193847656e-08, -0.5772156649015329, 0.5723649429247001, np.log(24.0)]),
    ("softplus <- function(v) { big <- v > 0; out <- log1p(exp(-abs(v))); out[big] <- out[big] + v[big]; out }; "
     "softplus(c(-800, 0, 800))", [0, np.log(2.0), 800]),
    ("f <- function() { x <- 1; g <- function() x <<- x + 1; g(); x }; f()", [2]),
    ("tryCatch({ stop('boom'); 1 }, error = function(e) 2)", [2]),
    ("uniroot(function(x) x^2 - 2, interval = c(0, 2), tol = 1e-12)$root", [2 ** 0.5]),
    ("m <- matrix(1:6, 2); dim(m[, 2, drop = FALSE])", [2, 1]),
    ("x <- c(3, 1, 2); rev(sort(x))", [3, 2, 1]),
    ("as.integer(0:(4 - 1))", [0, 1, 2, 3]),
    ("m <- matrix(1:4, 2); m[2, ] <- c(9, 8); as.vector(m)", [1, 9, 3, 8]),
    ("1:3 + 1:6", [2, 4, 6, 5, 7, 9]),
    ("max(5, c(1, 9)) - min(c(4, 2), 3)", [7]),
    ("!c(TRUE, FALSE) | c(FALSE, FALSE)", [0, 1]),
    ("x <- 5; if (x > 3 && x < 10) 1 else 0", [1]),
    ("length(NULL) + length(list(1, 2)) + nrow(matrix(0, 3, 2)) + ncol(matrix(0, 3, 2))", [7]),
]


@pytest.mark.parametrize("src,expected", R_SEMANTICS, ids=[s[0][:40] for s in R_SEMANTICS])
def test_r_semantics(src, expected):
    from oracle.rlite.interp import Interp
    v = Interp().run(src)
    got = np.asarray(v.a, dtype=np.float64).reshape(-1, order="F")
    np.testing.assert_allclose(got, np.asarray(expected, dtype=np.float64), rtol=1e-12, atol=1e-300, equal_nan=True)


def test_r_semantics_names_lists_and_match_call():
    from oracle.rlite.interp import Interp
    it = Interp()
    v = it.run("g <- function(...) setNames(list(...), as.character(match.call()[-1])); aa <- 1; bb <- 'x'; g(aa, bb)")
    assert v.names == ["aa", "bb"] and v.items[1].a[0] == "x"
    assert it.run("f2 <- function(x, eps = 2) deparse(substitute(x)); zz <- 3; f2(zz)").a[0] == "zz"
    assert list(it.run("paste0('a', 1:3, collapse = ', ')").a) == ["a1, a2, a3"]
    assert list(it.run("paste('x', 1.5)").a) == ["x 1.5"]
    assert list(it.run("names(c(a = 1, b = 2))").a) == ["a", "b"]
    m = it.run("m <- matrix(1:4, 2); colnames(m) <- c('u', 'v'); m[, c(FALSE, TRUE), drop = FALSE]")
    assert m.dimnames[1] == ["v"] and m.a.shape == (2, 1)
    with pytest.raises(Exception, match="must be positive"):
        it.run("chk <- function(x) if (any(x < 0)) stop(paste0(deparse(substitute(x)), ' must be positive')); tol <- -1; chk(tol)")
    assert it.run("x <- 1:4; x[c(TRUE, TRUE, FALSE, FALSE)] <- c(1, 2, 3); 1") is not None and it.warnings  # R warns too


# --------------------------------------------------------------------------------------------- golden vs restatement
@pytest.mark.parametrize("path", CORE_FILES, ids=CORE_IDS)
def test_restated_loop_reproduces_the_reference_r_code(oracle_built, path):
    from oracle import vb_oracle
    g, hyper, init, anneal = load_case(path)
    X, Y = g["X"], g["Y"]
    has_na = bool(np.isnan(Y).any())
    trace = []
    out = vb_oracle.atlasqtl_global_local_core_(Y, X, Y.shape[1], anneal, 1, float(g["tol"]), 1000, hyper, init,
                                                thinned_elbo_eval=bool(g["thinned"]), sweep="primal" if has_na else "dual",
                                                trace=trace)
    lb = np.array([r["lb"] for r in trace if r["lb"] is not None])
    assert out["converged"] and bool(g["converged"])
    assert out["it"] == int(g["it"])
    assert lb.shape == g["lb"].shape
    assert np.max(np.abs(lb - g["lb"]) / np.abs(g["lb"])) <= 1e-12
    assert abs(out["lb_opt"] - float(g["lb_opt"])) <= 1e-12 * abs(float(g["lb_opt"]))
    assert abs(out["diff_lb"] - float(g["diff_lb"])) <= 1e-8
    for k in ("gam_vb", "beta_vb", "theta_vb", "zeta_vb"):
        assert np.max(np.abs(out[k] - g[k])) <= 1e-10, k
    for k in ("tau_vb", "sig2_beta_vb", "sig2_theta_vb", "sig02_inv_vb", "lam2_inv_vb"):
        np.testing.assert_allclose(np.asarray(out[k], dtype=np.float64).reshape(g["full_" + k].shape), g["full_" + k],
                                   rtol=1e-9, err_msg=k)
    assert np.array_equal(out["gam_vb"] > 0.5, g["gam_vb"] > 0.5)


@pytest.mark.parametrize("path", CORE_FILES, ids=CORE_IDS)
def test_product_host_loop_reproduces_the_reference_r_code(oracle_built, path):
    """atlasqtl_b200.core (the loop the CUDA path runs under) with the oracle-backed test double for the device."""
    from atlasqtl_b200 import core
    from fake_context import OracleSweepContext
    g, hyper, init, anneal = load_case(path)
    if g["X"].shape[1] * g["Y"].shape[1] > 20000:
        pytest.skip("covered by the smaller cases; the test double is slow")
    X, Y = g["X"], g["Y"]
    trace = []
    out = core.atlasqtl_global_local_core_(Y, X, Y.shape[1], anneal, 1, float(g["tol"]), 1000, 0, hyper, init,
                                           full_output=True, thinned_elbo_eval=bool(g["thinned"]), debug=True, trace=trace,
                                           context_factory=lambda X_, Y_: OracleSweepContext(X_, Y_))
    lb = np.array([r["lb"] for r in trace if r["lb"] is not None])
    assert out["converged"] and out["it"] == int(g["it"])
    assert lb.shape == g["lb"].shape
    assert np.max(np.abs(lb - g["lb"]) / np.abs(g["lb"])) <= 1e-10
    for k in ("gam_vb", "beta_vb", "theta_vb", "zeta_vb"):
        assert np.max(np.abs(out[k] - g[k])) <= 1e-9, k
    for k in ("eta_vb", "kappa_vb", "lam2_inv_vb", "nu_s0_vb", "nu_vb", "rho_s0_vb", "rho_vb", "rho_xi_inv_vb",
              "sig02_inv_vb", "sig2_inv_vb", "sig2_theta_vb", "sig2_zeta_vb", "tau_vb", "xi_inv_vb"):
        np.testing.assert_allclose(np.asarray(out[k], dtype=np.float64).reshape(g["full_" + k].shape), g["full_" + k],
                                   rtol=1e-8, err_msg=k)


def test_c1_golden_trajectory_agrees_with_the_reference_r_code():
    """tests/golden/c1_trajectory.npz (made by oracle/vb_oracle.py, what the GPU C1 test compares with) against the same
    problem run through the reference's R code."""
    a = np.load(os.path.join(GOLD, "c1_trajectory.npz"))
    b = np.load(os.path.join(GOLD, "rlite_c1.npz"))
    np.testing.assert_allclose(a["in_check"], b["in_check"], rtol=1e-12)
    assert int(a["it"]) == int(b["it"]) and bool(a["converged"]) and bool(b["converged"])
    assert np.max(np.abs(a["lb"] - b["lb"]) / np.abs(b["lb"])) <= 1e-12
    assert np.array_equal(a["sel_ppi"], b["sel_ppi"])
    np.testing.assert_allclose(a["theta_vb"], b["theta_vb"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(a["zeta_vb"], b["zeta_vb"], rtol=1e-9, atol=1e-11)
    assert abs(float(a["sum_gam"]) - float(b["sum_gam"])) <= 1e-9


def test_c4_golden_trajectory_agrees_with_the_reference_r_code():
    """The same for BASELINE config C4 (n=500, p=10000, q=5000, 20 hotspots, anneal = c(1, 2, 10)) -- the config north_star
    designates for the ELBO-trajectory and PPI / bFDR selection parity check: c4_trajectory.npz (primal restatement,
    what the GPU C4 test compares with) against the run of the reference's own R code over its own dual-form loop, the
    {bFDR < 0.05} set formed by the reference's own assign_bFDR."""
    path = os.path.join(GOLD, "rlite_c4.npz")
    if not os.path.exists(path):
        pytest.skip("rlite_c4.npz has not been generated (make_rlite_golden.py c4, ~1.5 h)")
    a = np.load(os.path.join(GOLD, "c4_trajectory.npz"))
    b = np.load(path)
    np.testing.assert_allclose(a["in_check"], b["in_check"], rtol=1e-12)
    assert int(a["it"]) == int(b["it"]) and bool(a["converged"]) and bool(b["converged"])
    assert a["lb"].shape == b["lb"].shape
    assert np.max(np.abs(a["lb"] - b["lb"]) / np.abs(b["lb"])) <= 1e-11
    assert np.array_equal(a["sel_ppi"], b["sel_ppi"])
    assert np.array_equal(a["sel_fdr"], b["sel_fdr"])
    np.testing.assert_allclose(a["theta_vb"], b["theta_vb"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(a["zeta_vb"], b["zeta_vb"], rtol=1e-8, atol=1e-10)
    assert abs(float(a["sum_gam"]) - float(b["sum_gam"])) <= 1e-7
    common, ia, ib = np.intersect1d(a["probe_idx"], b["probe_idx"], return_indices=True)
    assert len(common) > 10000
    assert np.abs(a["probe_gam"][ia] - b["probe_gam"][ib]).max() <= 1e-9
    assert np.abs(a["probe_beta"][ia] - b["probe_beta"][ib]).max() <= 1e-9


# --------------------------------------------------------------------------------------------- functions
def test_function_level_fixtures(oracle_built):
    from atlasqtl_b200 import hyper_init, summarise
    from oracle import vb_oracle
    g = np.load(os.path.join(GOLD, "rlite_functions.npz"))
    for i, a in enumerate(g["ladder_args"]):
        np.testing.assert_allclose(vb_oracle.get_annealing_ladder_(tuple(a)), g[f"ladder_{i}"], rtol=1e-15)
    np.testing.assert_allclose(vb_oracle.Q_approx_vec(g["q_x"]), g["q_vec"], rtol=1e-14)
    # the scalar routine stops per element, the vector one when the slowest element has converged (R/utils.R:402):
    # they agree to the Lentz tolerance only, which is why the restatement must be the vector form
    assert 1e-12 < np.max(np.abs(g["q_scalar"] / g["q_vec"] - 1)) < 1e-6
    for i in range(4):
        with np.errstate(all="ignore"):
            mine = vb_oracle.update_annealed_lam2_inv_vb_(g["lam_L"], float(g[f"lam_c{i}"]), 1)
        np.testing.assert_allclose(mine, g[f"lam_{i}"], rtol=1e-13, equal_nan=True)
    U, lp, l1p = g["imr_U"], g["imr_logp"], g["imr_log1p"]
    np.testing.assert_allclose(vb_oracle.inv_mills_ratio_(1, U, l1p, lp), g["imr_1"], rtol=1e-15)
    np.testing.assert_allclose(vb_oracle.inv_mills_ratio_(0, U, l1p, lp), g["imr_0"], rtol=1e-15)
    np.testing.assert_allclose(vb_oracle.update_Z_(g["z_gam"], U, l1p, lp, 1.0), g["z_c1"], rtol=1e-15)
    np.testing.assert_allclose(vb_oracle.update_Z_(g["z_gam"], U, l1p, lp, 0.7), g["z_c07"], rtol=1e-14)
    # log(1 + exp(x)) of the sweep (src/coreLoop.cpp:28-33 is the C++ twin of R/utils.R:149-156)
    # -- written as the reference writes it: exact 0 below x = -36.7, where log1p-style routines return exp(x)
    x = g["l1pe_x"]
    m = np.maximum(x, 0)
    np.testing.assert_allclose(np.log(np.exp(x - m) + np.exp(-m)) + m, g["l1pe"], rtol=1e-15, atol=0)
    assert np.all(g["l1pe"][x < -37] == 0) and np.allclose(np.logaddexp(0, x), g["l1pe"], rtol=1e-12, atol=2e-16)
    for fn in (vb_oracle.assign_bFDR, summarise.assign_bFDR):
        np.testing.assert_allclose(fn(g["fdr_ppi"]), g["fdr"], rtol=1e-14)
        assert np.array_equal(fn(g["fdr_ppi"]) < 0.05, g["fdr"] < 0.05)
    for (E, V, p), o in zip(g["hyper_p0"], g["hyper_out"]):
        h = hyper_init.auto_set_hyper_(g["hyper_Y"], int(p), (E, V))
        # R's uniroot stops at its default tolerance .Machine$double.eps^0.25 = 1.2e-4 on t02; the mirror solves the
        # same equation to 1e-12, so the two agree to that tolerance only (root = "uniroot" reproduces R's iterate)
        assert abs(h["t02"] - o[0]) <= 2e-4
        assert abs(h["n0"][0] - o[1]) <= 1e-4 * abs(o[1])
        np.testing.assert_allclose([h["eta"][0], h["nu"], h["rho"]], o[2:], rtol=1e-13)
        hu = hyper_init.auto_set_hyper_(g["hyper_Y"], int(p), (E, V), root="uniroot")
        np.testing.assert_allclose([hu["t02"], hu["n0"][0]], o[:2], rtol=1e-10)


# --------------------------------------------------------------------------------------------- atlasqtl() end to end
def test_preprocessing_and_top_level_call_reproduce_the_reference_r_code(oracle_built):
    from atlasqtl_b200 import api, prepare
    from fake_context import OracleSweepContext
    from oracle import prepare_oracle
    g = np.load(os.path.join(GOLD, "rlite_atlasqtl_top.npz"))
    X, Y = g["X_raw"], g["Y_raw"]
    for dat in (prepare.prepare_data_(Y, X, 0.1, 1000), prepare_oracle.prepare_data_(Y, X)):
        assert np.array_equal(dat["bool_rmvd_x"], g["prep_bool_rmvd_x"])
        np.testing.assert_allclose(dat["X"], g["prep_X"], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(dat["Y"], g["prep_Y"], rtol=1e-13, atol=1e-13, equal_nan=True)
    dat = prepare.prepare_data_(Y, X, 0.1, 1000)
    assert dat["rmvd_cst_x"] == list(g["prep_rmvd_cst_x"])
    assert dat["initial_colnames_X"] == list(g["prep_initial_colnames_X"])
    pairs = sorted((k, r) for k, rs in dat["rmvd_coll_x"].items() for r in rs)
    assert pairs == sorted(zip(g["prep_rmvd_coll_kept"].tolist(), g["prep_rmvd_coll_x"].tolist()))
    assert dat["names_x"] == list(g["names_x"]) and dat["names_y"] == list(g["names_y"])
    hyper, init = hyper_init_of(g)
    trace = []
    out = api.atlasqtl(Y, X, None, anneal=tuple(g["anneal"]), tol=float(g["tol"]), maxit=1000, verbose=0,
                       list_hyper=hyper, list_init=init, trace=trace,
                       context_factory=lambda X_, Y_: OracleSweepContext(X_, Y_))
    lb = np.array([r["lb"] for r in trace if r["lb"] is not None])
    assert out["converged"] and out["it"] == int(g["it"])
    assert np.max(np.abs(lb - g["lb"]) / np.abs(g["lb"])) <= 1e-10
    for k in ("gam_vb", "beta_vb", "theta_vb", "zeta_vb"):
        assert np.max(np.abs(out[k] - g[k])) <= 1e-9, k
    # add_collinear_back = TRUE (add_collinear_back_, R/utils.R:680-735): rows for the removed duplicates again
    back = api.atlasqtl(Y, X, None, anneal=tuple(g["anneal"]), tol=float(g["tol"]), maxit=1000, verbose=0,
                        list_hyper=hyper, list_init=init, add_collinear_back=True,
                        context_factory=lambda X_, Y_: OracleSweepContext(X_, Y_))
    assert back["names_x"] == list(g["cb_names_x"]) and back["gam_vb"].shape == g["cb_gam_vb"].shape
    for k in ("gam_vb", "beta_vb", "theta_vb"):
        assert not np.isnan(back[k]).any()
        assert np.max(np.abs(back[k] - g["cb_" + k])) <= 1e-9, k
    rows = {nm: i for i, nm in enumerate(back["names_x"])}
    for kept, removed in zip(g["prep_rmvd_coll_kept"], g["prep_rmvd_coll_x"]):
        assert np.array_equal(back["gam_vb"][rows[removed]], back["gam_vb"][rows[kept]])


# --------------------------------------------------------------------------------------------- live re-derivation
def _live():
    from oracle.rlite import reference as R
    if not R.available():
        pytest.skip("/root/reference is not present here (GPU box): the committed fixtures stand in")
    return R


def test_fixtures_are_reproducible_from_the_reference_sources(oracle_built):
    R = _live()
    import make_rlite_golden as mk
    it = R.load()
    for name in ("b_geometric", "e_missing_anneal"):
        X, Y, hyper, init, anneal, thinned, tol = mk.core_case(name)
        g = np.load(os.path.join(GOLD, f"rlite_core_{name}.npz"))
        assert np.array_equal(X, g["X"]) and np.array_equal(Y, g["Y"], equal_nan=True)
        res = mk.run_core(it, X, Y, hyper, init, anneal, thinned, tol)
        assert res["it"] == int(g["it"])
        np.testing.assert_allclose(res["lb"], g["lb"], rtol=1e-13)
        np.testing.assert_allclose(res["gam_vb"], g["gam_vb"], rtol=0, atol=1e-12)
    assert not it.warnings


def test_reference_test_recipe_converges_through_its_own_r_code(oracle_built):
    """tests/testthat/main.R + test_convergence.R of the reference: n = 100, p = 75, q = 20, p0 = c(5, 25), default
    arguments, `expect_equal(vb$converged, TRUE)`.  R's RNG stream is not reproduced (inputs are drawn by NumPy and the
    starting values by the package's mirror), everything after that is the reference's atlasqtl()."""
    R = _live()
    from atlasqtl_b200 import hyper_init
    from oracle.rlite.values import from_py, to_py
    rng = np.random.default_rng(123)
    n, p, q, p_act = 100, 75, 20, 10
    X_act = rng.binomial(2, 0.2, size=(n, p_act)).astype(np.float64)
    X = np.concatenate([X_act, rng.binomial(2, 0.2, size=(n, p - p_act))], axis=1)[:, rng.permutation(p)]
    Y = X_act @ rng.normal(size=(p_act, q)) + rng.normal(size=(n, q))
    it = R.load()
    prep = it.call("prepare_data_", from_py(Y), from_py(X), from_py(0.1), from_py(1000.0), None, from_py(0.0), None, None)
    pp = prep.get("X").a.shape[1]
    init = {k: v for k, v in hyper_init.auto_set_init_(prep.get("Y").a, pp, (5, 25), q, user_seed=123).items()
            if not k.startswith("_")}
    vb = it.call("atlasqtl", Y=from_py(Y), X=from_py(X), p0=from_py(np.array([5.0, 25.0])), verbose=from_py(0.0),
                 list_init=R.with_class(R._copy_in(init), "out_init"))
    vb = to_py(vb)
    assert bool(vb["converged"][0])
    assert vb["gam_vb"].shape == (pp, q) and int(vb["it"][0]) < 1000
    # the product's atlasqtl() on the same call: default hyper-parameters (uniroot's iterate for t02, as R), default
    # annealing ladder, pre-processing, the loop -- against the reference's own atlasqtl()
    from atlasqtl_b200 import api
    from fake_context import OracleSweepContext
    out = api.atlasqtl(Y, X, (5.0, 25.0), verbose=0, list_init=init, save_hyper=True,
                       context_factory=lambda X_, Y_: OracleSweepContext(X_, Y_))
    assert out["converged"] and out["it"] == int(vb["it"][0])
    assert abs(out["lb_opt"] - float(vb["lb_opt"][0])) <= 1e-8 * abs(out["lb_opt"])
    assert np.abs(out["gam_vb"] - vb["gam_vb"]).max() <= 1e-7
    assert np.abs(out["beta_vb"] - vb["beta_vb"]).max() <= 1e-7


def test_reference_r_statement_of_the_sweep_equals_its_compiled_loop(oracle_built):
    """batch = "0" (R/atlasqtl_global_local_core.R:179-227) states the per-pair update in plain R -- pnorm(log.p),
    log_one_plus_exp_, the rank-1 correction of cp_X_Xbeta -- where batch = "y" calls the compiled coreDualLoop /
    coreDualMisLoop.  With `sample()` returning the identity order (the order coreDualLoop is given, :162-163) the two
    must walk the same trajectory: a check of the reference against itself, of the evaluator's handling of the in-place
    `.Call`, and of SciPy's log-CDF against the C++ loop's inputs."""
    R = _live()
    from oracle.rlite.values import Builtin, intv
    import make_rlite_golden as mk
    for name, size in (("b_geometric", (40, 18, 5)), ("e_missing_anneal", (40, 14, 4))):
        n, p, q = size
        X, Y, hyper, init, anneal, thinned, tol = mk.core_case(name)
        X, Y = np.asfortranarray(X[:n, :p]), np.asfortranarray(Y[:n, :q])
        hyper = dict(hyper, q_hyper=q, p_hyper=p, eta=hyper["eta"][:q], kappa=hyper["kappa"][:q], n0=hyper["n0"][:q])
        init = dict(init, q_init=q, p_init=p, gam_vb=init["gam_vb"][:p, :q], mu_beta_vb=init["mu_beta_vb"][:p, :q],
                    sig2_beta_vb=init["sig2_beta_vb"][:q], sig2_theta_vb=init["sig2_theta_vb"][:p],
                    tau_vb=init["tau_vb"][:q], theta_vb=init["theta_vb"][:p], zeta_vb=init["zeta_vb"][:q])
        it = R.load()
        it.globalenv.vars["sample"] = Builtin(lambda it_, pos, named: intv(pos[0].flat()), "sample")   # identity order
        runs = {}
        for batch in ("y", "0"):
            lbs = []
            out = R.global_local_core(Y, X, q, anneal, 1, tol, 200, hyper, init, it=it, hook=lambda k, v: lbs.append(v),
                                      batch=batch, debug=True)
            runs[batch] = (out, np.array(lbs))
        (a, la), (b, lb) = runs["y"], runs["0"]
        assert int(a["it"][0]) == int(b["it"][0]) and la.shape == lb.shape
        np.testing.assert_allclose(la, lb, rtol=1e-12)
        assert np.abs(a["gam_vb"] - b["gam_vb"]).max() <= 1e-11
        assert np.abs(a["beta_vb"] - b["beta_vb"]).max() <= 1e-11


def test_set_hyper_and_set_init_mirror_the_reference_constructors(oracle_built):
    """set_hyper / set_init (R/set_hyper_init.R:98-140, :311-351): same fields for valid arguments (scalars recycled
    to q), and the arguments the reference rejects are rejected."""
    R = _live()
    from atlasqtl_b200 import hyper_init
    from oracle.rlite.values import RError, from_py, to_py
    it = R.load()
    q, p = 4, 6
    rng = np.random.default_rng(2)

    def r_hyper(**kw):
        a = dict(q=float(q), p=float(p), eta=1.5, kappa=2.0, n0=-2.0, nu=0.01, rho=1.0, t02=0.2)
        a.update(kw)
        return it.call("set_hyper", *[from_py(a[k]) for k in ("q", "p", "eta", "kappa", "n0", "nu", "rho", "t02")])
    h_r = to_py(r_hyper())
    h_p = hyper_init.set_hyper(q, p, 1.5, 2.0, -2.0, 0.01, 1.0, 0.2)
    for k in ("A2_inv", "eta", "kappa", "m0", "n0", "nu", "rho", "t02", "q_hyper", "p_hyper"):
        np.testing.assert_allclose(np.asarray(h_p[k], dtype=float).reshape(-1), np.asarray(h_r[k], dtype=float).reshape(-1))
    h_r = to_py(r_hyper(eta=np.array([1.0, 2.0, 3.0, 4.0])))
    np.testing.assert_allclose(hyper_init.set_hyper(q, p, [1.0, 2.0, 3.0, 4.0], 2.0, -2.0, 0.01, 1.0, 0.2)["eta"], h_r["eta"])
    for bad in (dict(t02=-1.0), dict(nu=0.0), dict(eta=np.array([1.0, -1.0, 1.0, 1.0])), dict(kappa=np.array([1.0, 2.0]))):
        with pytest.raises(RError):
            r_hyper(**bad)
        with pytest.raises(ValueError):
            a = dict(eta=1.5, kappa=2.0, n0=-2.0, nu=0.01, rho=1.0, t02=0.2)
            a.update(bad)
            hyper_init.set_hyper(q, p, a["eta"], a["kappa"], a["n0"], a["nu"], a["rho"], a["t02"])

    good = dict(gam_vb=np.asfortranarray(rng.uniform(size=(p, q))), mu_beta_vb=np.asfortranarray(rng.normal(size=(p, q))),
                sig02_inv_vb=3.0, sig2_beta_vb=rng.uniform(0.5, 1, q), sig2_theta_vb=rng.uniform(0.5, 1, p),
                tau_vb=rng.uniform(0.5, 1, q), theta_vb=rng.normal(size=p), zeta_vb=rng.normal(size=q))
    order = ("gam_vb", "mu_beta_vb", "sig02_inv_vb", "sig2_beta_vb", "sig2_theta_vb", "tau_vb", "theta_vb", "zeta_vb")

    def r_init(**kw):
        a = dict(good)
        a.update(kw)
        return it.call("set_init", from_py(float(q)), from_py(float(p)), *[from_py(a[k]) for k in order])
    i_r = to_py(r_init())
    i_p = hyper_init.set_init(q, p, *[good[k] for k in order])
    for k in order + ("q_init", "p_init"):
        np.testing.assert_allclose(np.asarray(i_p[k], dtype=float), np.asarray(i_r[k], dtype=float).reshape(np.shape(i_p[k])))
    bad_gam = good["gam_vb"].copy()
    bad_gam[0, 0] = 1.5
    for bad in (dict(gam_vb=bad_gam), dict(tau_vb=-good["tau_vb"]), dict(theta_vb=good["theta_vb"][:-1]),
                dict(mu_beta_vb=good["mu_beta_vb"][:, :-1]), dict(sig02_inv_vb=-1.0)):
        with pytest.raises(RError):
            r_init(**bad)
        with pytest.raises(ValueError):
            a = dict(good)
            a.update(bad)
            hyper_init.set_init(q, p, *[a[k] for k in order])


def test_argument_checks_mirror_the_reference(oracle_built):
    """check_annealing_ (R/prepare_atlasqtl.R:101-128) and the data checks of prepare_data_ (:11-45): what the reference
    accepts is accepted, what it refuses is refused."""
    R = _live()
    from atlasqtl_b200 import core, prepare
    from oracle.rlite.values import RError, from_py
    it = R.load()
    for anneal, ok in (((1, 2, 10), True), ((2, 3, 7), True), ((3, 1.5, 2), True), (None, True), ((4, 2, 10), False),
                       ((1, 1.2, 10), False), ((1, 2, 1001), False), ((1, 2), False), ((1, 2, 2.5), False),
                       ((0, 2, 10), False), ((1, -2, 10), False)):
        arg = None if anneal is None else from_py(np.array(anneal, dtype=np.float64))
        if ok:
            it.call("check_annealing_", arg)
            core.check_annealing_(anneal)
        else:
            with pytest.raises(RError):
                it.call("check_annealing_", arg)
            with pytest.raises((ValueError, TypeError)):
                core.check_annealing_(anneal)
    rng = np.random.default_rng(4)
    X = rng.binomial(2, 0.3, size=(40, 12)).astype(np.float64)
    Y = rng.normal(size=(40, 5))
    few = np.full_like(Y, np.nan)
    few[:1, :] = 1.0                       # < 5 % non-NA overall
    col = Y.copy()
    col[:, 2] = np.nan                     # one column > 97.5 % NA
    edge = col.copy()
    edge[0, 2] = 0.3                       # exactly 2.5 % observed: `< 0.025` does not fire, in either
    it.call("prepare_data_", from_py(edge), from_py(X), from_py(0.1), from_py(10.0), None, from_py(0.0), None, None)
    prepare.prepare_data_(edge, X, 0.1, 10)
    for Yb, Xb, tol, maxit in ((few, X, 0.1, 10), (col[:, :], X, 0.1, 10), (Y[:-1], X, 0.1, 10), (Y, X, -1.0, 10),
                               (Y, X, 0.1, 2.5), (Y, np.ones_like(X), 0.1, 10)):
        with pytest.raises(RError):
            it.call("prepare_data_", from_py(Yb), from_py(Xb), from_py(float(tol)), from_py(float(maxit)), None,
                    from_py(0.0), None, None)
        with pytest.raises(ValueError):
            prepare.prepare_data_(Yb, Xb, tol, maxit)
