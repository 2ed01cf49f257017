"""GPU: golden vectors of the reference through the CUDA path, ragged / degenerate shapes, every kernel
configuration, error behaviour of the C ABI, and size-independent properties at a BASELINE-scale shape."""
import glob
import os

import numpy as np
import pytest

from problems import make_problem, sweep_inputs

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "coredualloop_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[14:-4] for p in GOLDEN])
def test_cuda_sweep_matches_reference_golden(path):
    """Outputs of the reference's own coreDualLoop (tests/golden/make_golden.py) vs the CUDA sweep."""
    from atlasqtl_b200.device import SweepContext
    d = dict(np.load(path))
    X, Y = np.asfortranarray(d["X"]), np.asfortranarray(d["Y"])
    with SweepContext(X, Y) as ctx:
        ctx.set_order(d["order"])
        ctx.set_state(d["gam"], d["mu"])
        ctx.refresh_tables(d["theta"], d["zeta"], c_next=float(d["c"]))
        ctx.sweep(float(d["c"]), float(d["log_sig2_inv"]), d["tau"], d["log_tau"], d["sig2_beta"])
        st = ctx.get_state()
        R = ctx.get_residual()
    assert np.abs(st["gam_vb"] - d["out_gam"]).max() <= 1e-9      # north_star bound is 1e-8
    assert np.abs(st["mu_beta_vb"] - d["out_mu"]).max() <= 1e-9
    assert np.abs(st["beta_vb"] - d["out_beta"]).max() <= 1e-9
    np.testing.assert_allclose(X.T @ (Y - R), d["out_cp_betaX_X"], atol=1e-8)  # the reference's running X'X beta


def _cpu_sweep(native, X, Y, si, order):
    gam, mu = si["gam"].copy(order="F"), si["mu"].copy(order="F")
    beta = np.asfortranarray(gam * mu)
    R = native.residual(X, Y, beta)
    native.sweep_primal(X, np.asfortranarray((X ** 2).sum(0)), R, gam, si["log_Phi"], si["log_1_min_Phi"],
                        si["log_sig2_inv"], si["log_tau"], beta, mu, si["sig2_beta"], si["tau"], order, c=si["c"],
                        nthreads=4)
    return gam, mu, R


@pytest.mark.parametrize("n,p,q", [
    (30, 1, 1),        # a single pair
    (30, 7, 3),        # p < one SNP block
    (64, 9, 65),       # p, q just past a block / tile boundary
    (144, 17, 33), (150, 40, 70), (216, 24, 64), (250, 33, 50), (360, 16, 48), (400, 20, 33),
    (504, 25, 32), (600, 30, 25), (720, 12, 24), (800, 21, 17), (1008, 10, 16),   # every single-CTA configuration
    # sample-split thread-block clusters (2 / 4 / 8 CTAs), every clustered configuration, several tiles per cluster
    (1009, 17, 40), (1100, 20, 70), (1250, 12, 30), (1400, 16, 20), (1500, 25, 50), (1728, 9, 17),
    (1800, 12, 40), (2500, 20, 33), (3000, 24, 50), (3456, 8, 16), (3600, 10, 30), (5000, 16, 60), (6912, 8, 24),
    # more tiles than one round of the persistent grid: full rounds by the main kernel + leftover traits by the 8-trait one
    (100, 16, 4776), (1000, 9, 2400), (1200, 12, 1806),
])
@pytest.mark.parametrize("tail", [True, False], ids=["tail8", "notail"])
def test_ragged_shapes_and_all_configs(oracle_built, monkeypatch, n, p, q, tail):
    """tail8: leftover traits (here usually ALL traits: fewer tiles than SMs) go through the 8-trait tile kernel;
    notail (AQ_NO_TAIL): every tile through the configuration's full-size kernel."""
    from atlasqtl_b200.device import SweepContext
    native = oracle_built
    if tail:
        monkeypatch.delenv("AQ_NO_TAIL", raising=False)
    else:
        monkeypatch.setenv("AQ_NO_TAIL", "1")
    rng = np.random.default_rng(n + p + q)
    X = rng.normal(size=(n, p))
    X = np.asfortranarray((X - X.mean(0)) / X.std(0, ddof=1))
    Y = rng.normal(size=(n, q))
    Y = np.asfortranarray(Y - Y.mean(0))
    si = sweep_inputs(X, Y, None, c=0.9)
    order = rng.permutation(p).astype(np.int32)
    g_ref, m_ref, R_ref = _cpu_sweep(native, X, Y, si, order)
    with SweepContext(X, Y) as ctx:
        ctx.set_order(order)
        ctx.set_state(si["gam"], si["mu"])
        ctx.refresh_tables(si["theta"], si["zeta"], c_next=0.9)
        out = ctx.sweep(0.9, si["log_sig2_inv"], si["tau"], si["log_tau"], si["sig2_beta"])
        st = ctx.get_state()
        R = ctx.get_residual()
    assert np.abs(st["gam_vb"] - g_ref).max() <= 1e-9
    assert np.abs(st["mu_beta_vb"] - m_ref).max() <= 1e-9
    np.testing.assert_allclose(R, R_ref, atol=1e-9)
    np.testing.assert_allclose(out["resid_sq"], (R_ref ** 2).sum(0), rtol=1e-10)


def test_two_sweeps_and_order_change(oracle_built):
    """State stays on the device between calls; changing shuffled_ind re-tiles X and changes the result."""
    from atlasqtl_b200.device import SweepContext
    native = oracle_built
    X, Y, hyper, init = make_problem(200, 120, 40)
    p = X.shape[1]
    si = sweep_inputs(X, Y, init, c=1.0)
    o1 = np.arange(p, dtype=np.int32)
    o2 = np.random.default_rng(3).permutation(p).astype(np.int32)
    gam, mu = si["gam"].copy(order="F"), si["mu"].copy(order="F")
    beta = np.asfortranarray(gam * mu)
    R = native.residual(X, Y, beta)
    xn = np.asfortranarray((X ** 2).sum(0))
    for o in (o1, o2):
        native.sweep_primal(X, xn, R, gam, si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"], si["log_tau"], beta, mu,
                            si["sig2_beta"], si["tau"], o, c=1.0)
    with SweepContext(X, Y) as ctx:
        ctx.set_state(si["gam"], si["mu"])
        ctx.refresh_tables(si["theta"], si["zeta"])
        for o in (None, o2):
            ctx.set_order(o)
            ctx.sweep(1.0, si["log_sig2_inv"], si["tau"], si["log_tau"], si["sig2_beta"])
        st = ctx.get_state()
    assert np.abs(st["gam_vb"] - gam).max() <= 1e-9


def test_error_behaviour():
    from atlasqtl_b200 import _lib
    from atlasqtl_b200.device import SweepContext
    rng = np.random.default_rng(0)
    X = np.asfortranarray(rng.normal(size=(40, 10)))
    Y = np.asfortranarray(rng.normal(size=(40, 5)))
    with SweepContext(X, Y) as ctx:
        with pytest.raises(_lib.AtlasqtlB200Error, match="before aq_set_state"):
            ctx.sweep(1.0, 0.0, np.ones(5), np.zeros(5), np.ones(5))
        ctx.set_state(np.full((10, 5), 0.1), np.zeros((10, 5)))
        with pytest.raises(_lib.AtlasqtlB200Error, match="before aq_refresh_tables"):
            ctx.sweep(1.0, 0.0, np.ones(5), np.zeros(5), np.ones(5))
        with pytest.raises(_lib.AtlasqtlB200Error, match="not a permutation"):
            ctx.set_order(np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 8], dtype=np.int32))
        with pytest.raises(_lib.AtlasqtlB200Error, match="not a permutation"):
            ctx.set_order(np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 10], dtype=np.int32))
        with pytest.raises(ValueError):
            ctx.set_state(np.zeros((9, 5)), np.zeros((10, 5)))
        ctx.refresh_tables(np.zeros(10), np.zeros(5))
        with pytest.raises(_lib.AtlasqtlB200Error, match="c must be positive"):
            ctx.sweep(0.0, 0.0, np.ones(5), np.zeros(5), np.ones(5))
        ctx.sweep(1.0, 0.0, np.ones(5), np.zeros(5), np.ones(5))  # the context is still usable after errors


def test_table_pass_matches_scipy_incl_tails():
    """D, W, I0 and the ELBO-B part against SciPy's log_ndtr (R's pnorm(log.p=TRUE)), deep tails included."""
    from scipy import special as sp

    from atlasqtl_b200.device import SweepContext
    rng = np.random.default_rng(4)
    p, q, n = 64, 40, 30
    X = np.asfortranarray(rng.normal(size=(n, p)))
    Y = np.asfortranarray(rng.normal(size=(n, q)))
    theta = np.linspace(-28, 9, p)
    zeta = rng.normal(-1.0, 2.0, q)
    gam = np.asfortranarray(rng.uniform(size=(p, q)) ** 6)
    gam[0, 0], gam[1, 1] = 0.0, 1.0
    u = theta[:, None] + zeta[None, :]
    eps = np.finfo(float).eps ** 0.75
    for c in (1.0, 0.5):
        with SweepContext(X, Y) as ctx:
            ctx.set_state(gam, np.zeros((p, q)))
            part = ctx.refresh_tables(theta, zeta, c_next=c, want_elbo=True)
            # the Z part is observable through the row sums: sum_k gam W + I0
            rows = ctx.rowsums_zpart()
        lp, lq = sp.log_ndtr(u), sp.log_ndtr(-u)
        expect = np.sum(gam * lp + (1 - gam) * lq - gam * np.log(gam + eps) - (1 - gam) * np.log(1 - gam + eps))
        assert abs(part - expect) <= 1e-12 * abs(expect)
        U = np.sqrt(c) * u
        lpU, lqU = sp.log_ndtr(U), sp.log_ndtr(-U)
        m1 = np.maximum(np.exp(-U ** 2 / 2 - 0.5 * np.log(2 * np.pi) - lpU), -U)
        m0 = np.minimum(-np.exp(-U ** 2 / 2 - 0.5 * np.log(2 * np.pi) - lqU), -U)
        np.testing.assert_allclose(rows, (gam * (m1 - m0) + m0).sum(axis=1), rtol=1e-11, atol=1e-10)


def test_size_independent_properties_at_scale():
    """BASELINE config C4 dimensions (n=500, p=10000, q=5000): no CPU oracle at this size, so check properties:
    (1) the residual the sweep carried equals Y - X beta rebuilt from scratch from the downloaded state;
    (2) a sweep at the fixed point of another sweep changes nothing it should not (column sums are consistent);
    (3) gam_vb stays in [0, 1] and finite;
    (4) the row sums of the Z part from the sweep's per-tile partials equal those of the streaming kernel."""
    import bench
    from atlasqtl_b200.device import SweepContext
    cfg, X, Y, hyper, init = bench.make_workload("C4", 0, 5000)
    n, p = X.shape
    q = Y.shape[1]
    tau = np.full(q, 1.0)
    sig2 = 1 / ((n - 1 + 1.0) * tau)
    with SweepContext(X, Y) as ctx:
        ctx.set_state(init["gam_vb"], init["mu_beta_vb"])
        ctx.refresh_tables(init["theta_vb"], init["zeta_vb"])
        out = ctx.sweep(1.0, 0.0, tau, np.zeros(q), sig2)
        rows_fused = ctx.rowsums_zpart()   # per-tile partials left by the sweep (+ one streamed row for the tail's traits)
        ctx.refresh_tables(init["theta_vb"], init["zeta_vb"])   # same tables again; drops the partials
        rows_streamed = ctx.rowsums_zpart()                     # streaming pass over gam, W, I0
        st = ctx.get_state()
        again = ctx.set_state(st["gam_vb"], st["mu_beta_vb"])  # rebuilds Y - X beta from scratch (mode 1)
    # (4) both routes to rowSums of the Z part agree, and add up to the column-sum route
    np.testing.assert_allclose(rows_fused, rows_streamed, rtol=1e-11, atol=1e-9)
    assert abs(rows_fused.sum() - out["colsum_zpart"].sum()) <= 1e-10 * np.abs(out["colsum_zpart"]).sum()
    assert np.isfinite(st["gam_vb"]).all() and st["gam_vb"].min() >= 0 and st["gam_vb"].max() <= 1
    np.testing.assert_allclose(out["resid_sq"], again["resid_sq"], rtol=1e-9)
    np.testing.assert_allclose(out["colsum_gam"], st["gam_vb"].sum(axis=0), rtol=1e-11)
    np.testing.assert_allclose(out["colsum_beta2"], (st["beta_vb"] ** 2).sum(axis=0), rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(out["colsum_gam"], again["colsum_gam"], rtol=1e-12)


@pytest.mark.parametrize("thres", [0.05, 0.25, 0.6])
def test_selection_sets_on_device_match_assign_bFDR(thres):
    """{PPI > thres} and {bFDR < thres} formed on the device (bisection on streaming count/sum passes) are exactly the
    sets of the reference's assign_bFDR / summary (R/summarise_output.R:99-106, :207-223), ties included."""
    from atlasqtl_b200 import summarise
    from atlasqtl_b200.device import SweepContext
    rng = np.random.default_rng(12)
    n, p, q = 40, 300, 130
    X = np.asfortranarray(rng.normal(size=(n, p)))
    Y = np.asfortranarray(rng.normal(size=(n, q)))
    gam = rng.uniform(size=(p, q)) ** 6
    hot = rng.uniform(size=(p, q)) < 0.03
    gam[hot] = 1 - rng.uniform(size=hot.sum()) ** 3 * 0.2
    gam[5:9, 3] = gam[11, 7] = gam[200, 100]          # exact ties
    gam[0, 0], gam[1, 1] = 1.0, 0.0
    gam = np.asfortranarray(gam)
    with SweepContext(X, Y) as ctx:
        ctx.set_state(gam, np.zeros((p, q)))
        rows, cols = summarise.select_ppi_device(ctx, thres)
        assert np.array_equal(np.stack([rows, cols], 1), np.argwhere(gam.T > thres)[:, ::-1])
        rows, cols, nsel = summarise.select_bFDR_device(ctx, thres)
    from oracle import vb_oracle
    want = vb_oracle.assign_bFDR(gam) < thres   # the oracle's restatement of R/summarise_output.R:207-223, not product code
    assert np.array_equal(want, summarise.assign_bFDR(gam) < thres)
    got = np.zeros_like(want)
    got[rows, cols] = True
    assert nsel == want.sum() == len(rows)
    assert np.array_equal(got, want)


def test_bfdr_device_ties_at_the_boundary():
    """All PPIs equal: the running mean is flat, so either everything or nothing is selected; and a block of ties that is
    only partly inside the prefix is cut in column-major order."""
    from atlasqtl_b200 import summarise
    from atlasqtl_b200.device import SweepContext
    rng = np.random.default_rng(1)
    n, p, q = 30, 16, 12
    X = np.asfortranarray(rng.normal(size=(n, p)))
    Y = np.asfortranarray(rng.normal(size=(n, q)))
    gam = np.full((p, q), 0.5)
    gam[:4, 0] = 0.99
    gam[2, 5] = gam[7, 2] = gam[9, 9] = gam[3, 1] = 0.8   # ties; with thres = 0.1 only some of them fit
    for thres in (0.1, 0.08, 0.3, 0.6):
        with SweepContext(X, Y) as ctx:
            ctx.set_state(np.asfortranarray(gam), np.zeros((p, q)))
            rows, cols, nsel = summarise.select_bFDR_device(ctx, thres)
        from oracle import vb_oracle
        want = vb_oracle.assign_bFDR(gam) < thres
        got = np.zeros_like(want)
        got[rows, cols] = True
        assert np.array_equal(got, want), thres


@pytest.mark.parametrize("margin", [1e-9, -1e-9, 1e-12, -1e-12])
def test_bfdr_device_near_threshold_boundary(margin):
    """Adversarial: thres is placed at (1 +- margin) x the running mean of 1 - PPI at one rank, so that the decision
    `cumsum(1 - ppi) / rank < thres` (R/summarise_output.R:216-218) of that element hangs on a relative 1e-9 / 1e-12.
    The reference decides with a sequential cumsum, the device path with fixed-order tree sums of the same terms: they can
    only disagree when |mean - thres| is at the rounding level of a sum of N terms (N = 2800 here: <= 3e-13 relative).
    Checked against an exactly rounded evaluation (math.fsum) and the ORACLE's assign_bFDR."""
    import math

    from atlasqtl_b200 import summarise
    from atlasqtl_b200.device import SweepContext
    from oracle import vb_oracle
    rng = np.random.default_rng(21)
    n, p, q = 30, 120, 40
    X = np.asfortranarray(rng.normal(size=(n, p)))
    Y = np.asfortranarray(rng.normal(size=(n, q)))
    gam = np.asfortranarray(1 - rng.uniform(size=(p, q)) ** 3)   # distinct with probability 1
    ee = np.sort(1 - gam.flatten())                                # what both sides see: e = 1 - gam_vb, ascending
    assert np.all(np.diff(ee) > 0)
    m = 2800
    mean_m1 = math.fsum(ee[:m + 1]) / (m + 1)
    thres = mean_m1 * (1 + margin)       # margin > 0: element m is the last one selected; margin < 0: the first one left out
    k_sel = m + 1 if margin > 0 else m
    exact_ok = lambda k: math.fsum(ee[:k]) / k < thres
    assert exact_ok(k_sel) and not exact_ok(k_sel + 1)
    with SweepContext(X, Y) as ctx:
        ctx.set_state(gam, np.zeros((p, q)))
        rows, cols, nsel = summarise.select_bFDR_device(ctx, thres)
    want = vb_oracle.assign_bFDR(gam) < thres
    assert want.sum() == k_sel
    got = np.zeros_like(want)
    got[rows, cols] = True
    assert nsel == k_sel
    assert np.array_equal(got, want)
