"""CPU: host-side mirror of the reference's R interface (argument checks, pre-processing, post-processing)."""
import numpy as np
import pytest

from atlasqtl_b200 import core, hyper_init, prepare, summarise
from problems import make_problem


def test_prepare_data_postconditions():
    rng = np.random.default_rng(0)
    G = rng.binomial(2, 0.3, size=(50, 12)).astype(float)
    G[:, 3] = 1.0          # constant column -> removed (R/prepare_atlasqtl.R:59)
    G[:, 7] = G[:, 2]      # duplicated column -> removed, first kept (:66)
    Y = rng.normal(size=(50, 4)) + 3.0
    d = prepare.prepare_data_(Y, G, 0.1, 10)
    X = d["X"]
    assert X.shape == (50, 10)
    np.testing.assert_allclose(X.sum(axis=0), 0, atol=1e-12)
    np.testing.assert_allclose((X ** 2).sum(axis=0), 49, rtol=1e-12)  # X_j'X_j = n - 1: the sweep's pre-condition
    np.testing.assert_allclose(d["Y"].mean(axis=0), 0, atol=1e-12)
    assert d["rmvd_cst_x"] == ["Cov_x_4"] and d["rmvd_coll_x"] == {"Cov_x_3": ["Cov_x_8"]}
    assert d["bool_rmvd_x"].sum() == 2


def test_annealing_checks_and_ladder():
    core.check_annealing_(None)
    core.check_annealing_((1, 2, 10))
    for bad in ((4, 2, 10), (1, 1.2, 10), (1, 2, 2000), (1, 2)):
        with pytest.raises(ValueError):
            core.check_annealing_(bad)
    lad = core.get_annealing_ladder_((1, 2, 10))
    np.testing.assert_allclose(lad[[0, -1]], [0.5, 1.0])          # R/utils.R:113-122
    np.testing.assert_allclose(lad[1:] / lad[:-1], 2 ** (1 / 9))
    np.testing.assert_allclose(core.get_annealing_ladder_((3, 2, 5)), np.linspace(0.5, 1, 5))
    np.testing.assert_allclose(core.get_annealing_ladder_((2, 2, 3)), [1 / 2, 1 / 1.5, 1.0])


def test_set_hyper_set_init_validation():
    h = hyper_init.set_hyper(4, 6, 1.0, 1.0, -2.0, 0.01, 1.0, 0.3)
    assert h["eta"].shape == (4,) and h["A2_inv"] == 1.0 and h["m0"] == 0.0
    with pytest.raises(ValueError):
        hyper_init.set_hyper(4, 6, -1.0, 1.0, -2.0, 0.01, 1.0, 0.3)
    g = np.full((6, 4), 0.5)
    ok = hyper_init.set_init(4, 6, g, g, 1.0, np.ones(4), np.ones(6), np.ones(4), np.zeros(6), np.zeros(4))
    assert ok["p_init"] == 6
    with pytest.raises(ValueError):
        hyper_init.set_init(4, 6, g * 3, g, 1.0, np.ones(4), np.ones(6), np.ones(4), np.zeros(6), np.zeros(4))
    with pytest.raises(ValueError):
        hyper_init.set_init(4, 6, g, g, 1.0, np.ones(3), np.ones(6), np.ones(4), np.zeros(6), np.zeros(4))


def test_auto_hyper_matches_requested_prior_moments():
    X, Y, hyper, init = make_problem(80, 60, 10, p0=(4, 12))
    p = X.shape[1]
    mu = hyper_init.get_mu(4, hyper["t02"], p)
    np.testing.assert_allclose(hyper["n0"], mu)
    np.testing.assert_allclose(hyper_init.get_V_p_t(mu, hyper["t02"], p), 12, rtol=1e-6)
    with pytest.raises(ValueError):
        hyper_init._solve_t02(50000, (100.0, 10.0))  # variance below the binomial floor: uniroot fails in R too


def test_assign_bfdr_matches_definition_and_ties_are_stable():
    ppi = np.array([[0.9, 0.5, 0.2], [0.99, 0.5, 0.6]])
    fdr = summarise.assign_bFDR(ppi)
    vec = ppi.flatten(order="F")
    for idx in range(vec.size):
        order = sorted(range(vec.size), key=lambda i: (-vec[i], i))
        rank = order.index(idx) + 1
        expect = sum(1 - vec[i] for i in order[:rank]) / rank
        assert abs(fdr.flatten(order="F")[idx] - expect) < 1e-15
    sel = summarise.selected_pairs(ppi, 0.55)
    assert {tuple(x) for x in sel} == {(0, 0), (1, 0), (1, 2)}


def test_pack_genotypes_layout():
    """Packed calls as aq_prep_geno reads them: sample i of a column = bits 2 (i % 4) .. 2 (i % 4) + 1 of byte i / 4."""
    from atlasqtl_b200.device import pack_genotypes
    rng = np.random.default_rng(1)
    for n in (1, 4, 7, 50):
        G = rng.integers(0, 3, size=(n, 9))
        g = pack_genotypes(G)
        assert g.dtype == np.uint8 and g.shape == (9, (n + 3) // 4)
        back = np.stack([(g[:, i // 4] >> (2 * (i % 4))) & 3 for i in range(n)])
        assert np.array_equal(back, G)
        if n % 4:  # padding samples of the last byte are call 0
            assert ((g[:, -1] >> (2 * (n % 4))) == 0).all()
    for bad in (np.array([[0, 3]]), np.array([[0.5, 1.0]]), np.array([[-1, 0]])):
        with pytest.raises(ValueError):
            pack_genotypes(bad)


def test_prepare_data_device_checks_before_touching_the_gpu():
    """Argument checks of prepare_data_ (R/prepare_atlasqtl.R:11-45) come first; without a GPU the device path then
    raises instead of falling back to the host implementation."""
    import torch

    from atlasqtl_b200 import _lib
    rng = np.random.default_rng(2)
    G = rng.binomial(2, 0.3, size=(50, 6)).astype(float)
    Y = rng.normal(size=(50, 3))
    with pytest.raises(ValueError, match="same number of samples"):
        prepare.prepare_data_device_(Y[:30], G, 0.1, 10)
    with pytest.raises(ValueError, match="tol"):
        prepare.prepare_data_device_(Y, G, 0.0, 10)
    Yn = Y.copy()
    Yn[1:, 1] = np.nan
    with pytest.raises(ValueError, match="97.5% missing"):
        prepare.prepare_data_device_(Yn, G, 0.1, 10)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.AtlasqtlB200Error):
            prepare.prepare_data_device_(Y, G, 0.1, 10)


def test_hyper_init_ignore_missing_responses():
    """eta / tau come from apply(Y, 2, var, na.rm = TRUE) (R/set_hyper_init.R:153)."""
    rng = np.random.default_rng(3)
    Y = rng.normal(size=(60, 5))
    Yn = Y.copy()
    Yn[::7, 2] = np.nan
    h = hyper_init.auto_set_hyper_(Yn, 20, (2, 10))
    assert np.isfinite(h["eta"]).all()
    expect = 1 / np.median([np.var(Yn[~np.isnan(Yn[:, k]), k], ddof=1) for k in range(5)])
    np.testing.assert_allclose(np.atleast_1d(h["eta"])[0], expect, rtol=1e-12)
