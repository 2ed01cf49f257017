"""GPU: pre-processing on the device (aq_prep_x / aq_prep_geno / aq_create_prepared) against the CPU oracle's literal
restatement of prepare_data_ (oracle/prepare_oracle.py; R/prepare_atlasqtl.R:57-83, rm_constant_ / rm_collinear_
R/utils.R:276-343) and against the package's host mirror of the same function.

Integer outputs (which columns are constant / duplicated, and of which kept column) must match exactly; the
standardised X and the centred Y to 1e-13 absolute (values are O(1); the summation order of a mean differs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-13


def raw_problem(n, p, q, seed=5, continuous=True, na_frac=0.0):
    rng = np.random.default_rng(seed)
    G = rng.binomial(2, 0.25, size=(n, p)).astype(np.float64)
    if continuous:  # some non-genotype predictors
        G[:, ::7] = rng.normal(3.0, 2.0, size=(n, len(range(0, p, 7))))
    # constant columns (all-0, all-2, a constant double), exact duplicates (pairs, a triple, a duplicate of a duplicate's source)
    G[:, 3] = 0.0
    G[:, 11] = 2.0
    if continuous:
        G[:, 14] = 0.25
    G[:, 20] = G[:, 5]
    G[:, 21] = G[:, 5]
    G[:, p - 1] = G[:, 8]
    G[:, 30] = G[:, p - 2]      # the LATER column is the original here: column 30 is kept, p - 2 dropped
    G[:, 40] = 2.0 - G[:, 9]    # sign-flipped coding is not a duplicate
    beta = np.zeros((p, q))
    beta[rng.choice(p, 5, replace=False)] = rng.normal(size=(5, q))
    Y = G @ beta + rng.normal(size=(n, q)) + 10.0
    if na_frac:
        Y[rng.uniform(size=Y.shape) < na_frac] = np.nan
        Y[:, ::4] = np.where(np.isnan(Y[:, ::4]), 1.0, Y[:, ::4])
    return np.asfortranarray(G), np.asfortranarray(Y)


def check_against_oracle(G, Y_raw, prep, ctx):
    """Against the literal restatement of R's scale / rm_constant_ / rm_collinear_ (oracle/prepare_oracle.py)."""
    from oracle import prepare_oracle
    o = prepare_oracle.prepare_data_(Y_raw, G)
    assert np.array_equal(prep.status == 1, o["bool_cst_x"])
    assert np.array_equal(prep.status != 0, o["bool_rmvd_x"])
    assert np.array_equal(np.flatnonzero(prep.status == 0), o["kept"])
    assert np.array_equal(np.where(prep.status == 2, prep.dup_of, -1), o["dup_of"])
    assert np.abs(ctx.get_x() - o["X"]).max() <= TOL
    assert np.abs(ctx.get_y() - np.where(np.isnan(o["Y"]), 0.0, o["Y"])).max() <= TOL


def check_against_host(dat_h, prep, ctx, Y_raw):
    assert prep.p == dat_h["X"].shape[1]
    assert np.array_equal(prep.status != 0, dat_h["bool_rmvd_x"])
    names = [f"Cov_x_{j + 1}" for j in range(prep.p_raw)]
    assert [names[j] for j in np.flatnonzero(prep.status == 1)] == dat_h["rmvd_cst_x"]
    coll = {}
    for j in np.flatnonzero(prep.status == 2):
        coll.setdefault(names[prep.dup_of[j]], []).append(names[j])
    assert coll == dat_h["rmvd_coll_x"]
    Xd = ctx.get_x()
    assert np.abs(Xd - dat_h["X"]).max() <= TOL
    Yd = ctx.get_y()
    Yh = np.where(np.isnan(dat_h["Y"]), 0.0, dat_h["Y"])
    assert np.abs(Yd - Yh).max() <= TOL
    assert np.array_equal(ctx.n_obs, (~np.isnan(Y_raw)).sum(axis=0).astype(np.float64))
    # post-conditions the sweep relies on: zero column means, X_j'X_j = n - 1
    n = Xd.shape[0]
    assert np.abs(Xd.sum(axis=0)).max() <= 1e-10
    assert np.abs((Xd ** 2).sum(axis=0) - (n - 1)).max() <= 1e-9


@pytest.mark.parametrize("n,p,q,na", [(200, 120, 17, 0.0), (333, 257, 40, 0.1), (1001, 64, 9, 0.0)])
def test_prep_doubles_matches_host(n, p, q, na):
    from atlasqtl_b200 import device, prepare
    G, Y = raw_problem(n, p, q, na_frac=na)
    dat_h = prepare.prepare_data_(Y, G, 0.1, 10)
    with_prep = device.PreparedPredictors(G)
    with with_prep.context(Y) as ctx:
        check_against_host(dat_h, with_prep, ctx, Y)
        check_against_oracle(G, Y, with_prep, ctx)
    assert with_prep.launch_count() >= 2  # moments + duplicate verification ran on the device
    with_prep.close()


@pytest.mark.parametrize("n,p,q", [(200, 120, 17), (203, 300, 8), (1001, 64, 9)])
def test_prep_packed_genotypes_matches_host(n, p, q):
    from atlasqtl_b200 import device, prepare
    G, Y = raw_problem(n, p, q, continuous=False)
    dat_h = prepare.prepare_data_(Y, G, 0.1, 10)
    packed = device.pack_genotypes(G)
    assert packed.shape == (p, (n + 3) // 4)
    prep = device.PreparedPredictors(packed=packed, n=n)
    with prep.context(Y) as ctx:
        check_against_host(dat_h, prep, ctx, Y)
        check_against_oracle(G, Y, prep, ctx)
    # padded column stride
    wide = np.zeros((p, packed.shape[1] + 5), np.uint8)
    wide[:, :packed.shape[1]] = packed
    prep2 = device.PreparedPredictors(packed=wide, n=n)
    assert np.array_equal(prep2.status, prep.status) and np.array_equal(prep2.dup_of, prep.dup_of)
    assert np.array_equal(prep2.sd, prep.sd)
    prep.close()
    prep2.close()


def test_prep_errors():
    from atlasqtl_b200 import _lib, device
    G, Y = raw_problem(100, 60, 5, continuous=False)
    packed = device.pack_genotypes(G)
    bad = packed.copy()
    bad[17, 3] |= 0xC0  # call code 3
    with pytest.raises(_lib.AtlasqtlB200Error, match="column 17"):
        device.PreparedPredictors(packed=bad, n=100)
    Gn = G.copy()
    Gn[5, 9] = np.nan
    with pytest.raises(_lib.AtlasqtlB200Error, match="column 9"):
        device.PreparedPredictors(Gn)
    with pytest.raises(ValueError):
        device.PreparedPredictors(packed=packed[:, :10], n=100)  # stride shorter than ceil(n / 4)
    allc = np.ones((50, 4))
    prep = device.PreparedPredictors(allc)
    assert prep.p == 0 and (prep.status == 1).all()
    with pytest.raises(_lib.AtlasqtlB200Error, match="non-constant"):
        prep.context(np.zeros((50, 2)))
    # a trait without any observed value
    prep = device.PreparedPredictors(G)
    Yn = Y.copy()
    Yn[:, 2] = np.nan
    with pytest.raises(_lib.AtlasqtlB200Error, match="no observed value"):
        prep.context(Yn)


def test_many_duplicates_and_groups():
    """Every column appears three times in shuffled positions: 2/3 are dropped, each in favour of its first copy."""
    from atlasqtl_b200 import device
    rng = np.random.default_rng(3)
    n, p0 = 150, 500
    base = rng.binomial(2, 0.3, size=(n, p0)).astype(np.float64)
    base[0, :] = 0.0
    base[1, :] = 1.0  # no constant column
    src = rng.permutation(np.repeat(np.arange(p0), 3))
    G = np.asfortranarray(base[:, src])
    prep = device.PreparedPredictors(packed=device.pack_genotypes(G), n=n)
    first = {}
    for j, s in enumerate(src):
        first.setdefault(tuple(base[:, s]), j)
    exp_dup = np.array([first[tuple(base[:, s])] for s in src])
    exp_status = np.where(exp_dup == np.arange(3 * p0), 0, 2)
    assert np.array_equal(prep.status, exp_status)
    assert np.array_equal(prep.dup_of[exp_status == 2], exp_dup[exp_status == 2])
    assert (prep.dup_of[exp_status == 0] == -1).all()
    prep.close()


@pytest.mark.parametrize("packed,na", [(False, 0.0), (True, 0.0), (True, 0.06)])
def test_full_run_device_prepare_equals_host_prepare(packed, na):
    """atlasqtl() end to end: prepare_data_ on the device (doubles / packed calls, with and without missing responses)
    reproduces the run on host-prepared data: same iterations, ELBO to 1e-10 relative, gam_vb to 1e-8, same selections."""
    from atlasqtl_b200 import atlasqtl, device
    G, Y = raw_problem(150, 90, 24, seed=9, continuous=not packed, na_frac=na)
    kw = dict(p0=(3, 10), anneal=(1, 2, 5), tol=0.1, maxit=200, user_seed=123, verbose=0)
    tr_h, tr_d = [], []
    ref = atlasqtl(Y, G, trace=tr_h, **kw)
    if packed:
        out = atlasqtl(Y, device.pack_genotypes(G), packed_n=G.shape[0], trace=tr_d, **kw)
    else:
        out = atlasqtl(Y, G, prepare_on_device=True, trace=tr_d, **kw)
    assert out["converged"] and out["it"] == ref["it"]
    assert out["rmvd_cst_x"] == ref["rmvd_cst_x"] and out["rmvd_coll_x"] == ref["rmvd_coll_x"]
    assert out["names_x"] == ref["names_x"]
    for a, b in zip(tr_h, tr_d):
        if a["lb"] is not None:
            assert abs(a["lb"] - b["lb"]) <= 1e-10 * abs(a["lb"])
    assert np.abs(out["gam_vb"] - ref["gam_vb"]).max() <= 1e-8
    assert np.array_equal(out["gam_vb"] > 0.5, ref["gam_vb"] > 0.5)
