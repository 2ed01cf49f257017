"""`atlasqtl()`: the reference's user entry point (R/atlasqtl.R:179-322) on top of the CUDA hot path.

Same arguments, same output object (a dict with the fields of the R list, R/atlasqtl_global_local_core.R:414-428
plus p0 / rmvd_cst_x / rmvd_coll_x, R/atlasqtl.R:293-314).  Extra keyword-only arguments select the GPU
and, under torch.distributed, the slab of traits this process owns.
"""
import numpy as np

from . import core, hyper_init, prepare


def add_collinear_back_(beta_vb, gam_vb, theta_vb, names_x, initial_colnames_X, rmvd_coll_x):
    """R/utils.R:680-735: the posterior summaries with one row per NON-CONSTANT predictor again -- a predictor removed as
    a duplicate gets the rows of the predictor it duplicates.  rmvd_coll_x: {kept name: [removed names]}
    (names(rmvd_coll_x) -> rmvd_coll_x in R).  Returns beta_vb, gam_vb, theta_vb and names_x over initial_colnames_X."""
    pos = {nm: i for i, nm in enumerate(initial_colnames_X)}
    kept_rows = np.array([pos[nm] for nm in names_x])
    p_all, q = len(initial_colnames_X), gam_vb.shape[1]
    out = {}
    for key, arr in (("beta_vb", beta_vb), ("gam_vb", gam_vb)):
        full = np.full((p_all, q), np.nan, order="F")
        full[kept_rows] = arr
        for kept, removed in rmvd_coll_x.items():
            for nm in removed:
                full[pos[nm]] = full[pos[kept]]
        out[key] = full
    th = np.full(p_all, np.nan)
    th[kept_rows] = theta_vb
    for kept, removed in rmvd_coll_x.items():
        for nm in removed:
            th[pos[nm]] = th[pos[kept]]
    out["theta_vb"] = th
    out["names_x"] = list(initial_colnames_X)
    return out


def atlasqtl(Y, X, p0, anneal=(1, 2, 10), tol=0.1, maxit=1000, user_seed=None, verbose=1, list_hyper=None,
             list_init=None, save_hyper=False, save_init=False, full_output=False, thinned_elbo_eval=True,
             checkpoint_path=None, trace_path=None, add_collinear_back=False, *, device=0, comm=None, order_fn=None,
             trace=None, context_factory=None, prepare_on_device=False, packed_n=None, hyper_root="uniroot"):
    """hyper_root: how the default hyper-parameters / starting values solve for t02 -- "uniroot" stops where R's uniroot
    does (its default tolerance), reproducing the reference's t02 / n0; "brentq" solves the equation to 1e-12.
    prepare_on_device=True runs prepare_data_'s X / Y work on the GPU (X standardised there, never materialised on
    the host); with packed_n = n, X holds packed 2-bit genotype calls (`device.pack_genotypes`) instead of doubles."""
    if verbose not in (0, 1, 2):
        raise ValueError("The verbose argument must be set to 0, 1 or 2.")
    core.check_annealing_(anneal)
    if prepare_on_device or packed_n is not None:
        dat = prepare.prepare_data_device_(Y, X, tol, maxit, user_seed, verbose, checkpoint_path, packed_n=packed_n,
                                           device=device)
    else:
        dat = prepare.prepare_data_(Y, X, tol, maxit, user_seed, verbose, checkpoint_path)
    Xp, Yp = dat["X"], dat["Y"]  # device path: Xp is a handle with .shape, Yp the raw responses (hyper / init use variances only)
    n, p = Xp.shape
    q = Yp.shape[1]
    shr_fac_inv = q  # R/atlasqtl.R:218
    if list_hyper is None or list_init is None:
        if p0 is None or len(p0) != 2 or min(p0) <= 0:
            raise ValueError("p0 must be a vector of two positive numbers.")
    if list_hyper is None:
        list_hyper = hyper_init.auto_set_hyper_(Yp, p, p0, root=hyper_root)
    elif list_hyper["p_hyper"] != p or list_hyper["q_hyper"] != q:
        raise ValueError("The dimensions of list_hyper do not match those of the (pre-processed) data.")
    if list_init is None:
        if comm is not None and comm.world_size > 1 and user_seed is None:
            # theta_vb, sig2_theta_vb, sig02_inv_vb are REPLICATED state: every rank must draw the same initial values, so
            # the seed of an unseeded run is drawn once, on rank 0, and shared (all-reduce of [seed, 0, 0, ...])
            seed0 = float(np.random.SeedSequence().entropy % (1 << 52)) if comm.rank == 0 else 0.0
            user_seed = int(comm.allreduce_sum(np.array([seed0]))[0])
        list_init = hyper_init.auto_set_init_(Yp, p, p0, shr_fac_inv, user_seed, root=hyper_root)
    elif list_init["p_init"] != p or list_init["q_init"] != q:
        raise ValueError("The dimensions of list_init do not match those of the (pre-processed) data.")
    if add_collinear_back and comm is not None and comm.world_size > 1:
        raise NotImplementedError("add_collinear_back under a communicator: apply add_collinear_back_ to the gathered result")

    df = 1  # R/atlasqtl.R:272 (hs <- TRUE, debug <- TRUE)
    slab = None
    Y_loc = Yp
    if comm is not None and comm.world_size > 1:
        from .dist import slab_bounds
        slab = slab_bounds(q, comm.rank, comm.world_size)
        Y_loc = np.asfortranarray(Yp[:, slab[0]:slab[1]])
    res = core.atlasqtl_global_local_core_(Y_loc, Xp, shr_fac_inv, anneal, df, tol, maxit, verbose, list_hyper,
                                           list_init, checkpoint_path, trace_path, full_output, thinned_elbo_eval,
                                           debug=True, comm=comm, slab=slab, device=device, order_fn=order_fn,
                                           trace=trace, context_factory=context_factory)
    if hasattr(Xp, "close"):
        Xp.close()  # device path: the raw predictors held for aq_create_prepared are no longer needed
    if comm is not None and comm.world_size > 1:
        res = comm.gather_result(res, q)
    res["p0"] = p0
    res["rmvd_cst_x"] = dat["rmvd_cst_x"]
    res["rmvd_coll_x"] = dat["rmvd_coll_x"]
    res["names_x"], res["names_y"] = dat["names_x"], dat["names_y"]
    if add_collinear_back and len(dat["rmvd_coll_x"]) > 0:  # R/atlasqtl.R:297-310
        res.update(add_collinear_back_(res["beta_vb"], res["gam_vb"], res["theta_vb"], dat["names_x"],
                                       dat["initial_colnames_X"], dat["rmvd_coll_x"]))
    if save_hyper:
        res["list_hyper"] = list_hyper
    if save_init:
        res["list_init"] = list_init
    res["_class"] = "atlasqtl"
    return res
