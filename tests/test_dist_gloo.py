"""CPU, world_size = 2 over gloo: the trait-slab sharding and the all-reduce protocol of atlasqtl_b200.core /
dist reproduce the single-process run (the CUDA context is replaced by the oracle-backed test double)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from problems import make_problem

HERE = os.path.dirname(os.path.abspath(__file__))


def _with_nans(Y, frac):
    if not frac:
        return Y
    Ym = Y.copy()
    Ym[np.random.default_rng(3).uniform(size=Y.shape) < frac] = np.nan
    return Ym


def _worker(rank, world, port, tmpdir, anneal, nan_frac=0.0, golden=None):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    from atlasqtl_b200 import core
    from atlasqtl_b200.dist import TorchComm, slab_bounds
    from fake_context import OracleSweepContext
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    if golden is not None:   # inputs of a fixture made by the reference's own R code (tests/golden/rlite_core_*.npz)
        from rlite_cases import load_case
        g, hyper, init, anneal = load_case(golden)
        X, Y = g["X"], g["Y"]
    else:
        X, Y, hyper, init = make_problem(100, 75, 21, p_act=10, q_act=20, maf=0.2, p0=(5, 25))
        Y = _with_nans(Y, nan_frac)
    q = Y.shape[1]
    k0, k1 = slab_bounds(q, rank, world)
    comm = TorchComm()
    tr = []
    out = core.atlasqtl_global_local_core_(np.asfortranarray(Y[:, k0:k1]), X, q, anneal, 1, 0.1, 1000, 0, hyper, init,
                                           debug=True, comm=comm, slab=(k0, k1), trace=tr,
                                           context_factory=lambda X_, Y_: OracleSweepContext(X_, Y_))
    full = comm.gather_result(out, q)
    if rank == 0:
        np.savez(os.path.join(tmpdir, "dist.npz"), gam=full["gam_vb"], zeta=full["zeta_vb"], theta=full["theta_vb"],
                 it=out["it"], lbs=np.array([r["lb"] for r in tr if r["lb"] is not None]))
    dist.destroy_process_group()


def test_slab_bounds_cover_all_traits():
    from atlasqtl_b200.dist import slab_bounds
    for q, w in ((20, 8), (21, 2), (5, 8), (20000, 8)):
        b = [slab_bounds(q, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == q
        assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert max(x[1] - x[0] for x in b) - min(x[1] - x[0] for x in b) <= 1


@pytest.mark.parametrize("anneal,nan_frac", [((1, 2, 10), 0.0), (None, 0.0), ((1, 2, 5), 0.06)])
def test_two_slabs_reproduce_single_process(oracle_built, tmp_path, anneal, nan_frac):
    """nan_frac > 0: the missing-response path (per-slab mis_pat / X_norm_sq, colSums(mis_pat) in eta_vb and e_y_)."""
    from atlasqtl_b200 import core
    from fake_context import OracleSweepContext
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path), anneal, nan_frac), nprocs=2, join=True)
    d = np.load(tmp_path / "dist.npz")
    X, Y, hyper, init = make_problem(100, 75, 21, p_act=10, q_act=20, maf=0.2, p0=(5, 25))
    Y = _with_nans(Y, nan_frac)
    tr = []
    one = core.atlasqtl_global_local_core_(Y, X, Y.shape[1], anneal, 1, 0.1, 1000, 0, hyper, init, debug=True, trace=tr,
                                           context_factory=lambda X_, Y_: OracleSweepContext(X_, Y_))
    assert int(d["it"]) == one["it"]
    lbs = np.array([r["lb"] for r in tr if r["lb"] is not None])
    np.testing.assert_allclose(d["lbs"], lbs, rtol=1e-11)
    assert np.abs(d["gam"] - one["gam_vb"]).max() <= 1e-10
    np.testing.assert_allclose(d["theta"], one["theta_vb"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(d["zeta"], one["zeta_vb"], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("case", ["b_geometric", "e_missing_anneal"])
def test_two_slabs_reproduce_the_reference_r_code(oracle_built, tmp_path, case):
    """The sharded run (2 ranks over gloo, one all-reduce per sweep) against outputs of the reference's own R code."""
    from rlite_cases import GOLD, load_case
    path = os.path.join(GOLD, f"rlite_core_{case}.npz")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path), None, 0.0, path), nprocs=2, join=True)
    d = np.load(tmp_path / "dist.npz")
    g = load_case(path)[0]
    assert int(d["it"]) == int(g["it"])
    np.testing.assert_allclose(d["lbs"], g["lb"], rtol=1e-10)
    assert np.abs(d["gam"] - g["gam_vb"]).max() <= 1e-9
    np.testing.assert_allclose(d["theta"], g["theta_vb"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(d["zeta"], g["zeta_vb"], rtol=1e-8, atol=1e-10)


class _PpiContext:
    """Test double of the selection entry points of SweepContext over a host matrix (one slab)."""

    def __init__(self, gam):
        self.gam = np.asarray(gam)
        self.p, self.q = self.gam.shape

    def ppi_count_sum(self, t):
        e = 1.0 - self.gam
        m = e <= t
        return float(m.sum()), float(e[m].sum())

    def ppi_next_above(self, t):
        e = 1.0 - self.gam
        return float(e[e > t].min()) if (e > t).any() else float("inf")

    def ppi_collect(self, mode, lo, hi, capacity):
        e = 1.0 - self.gam
        m = (e > lo) & (e <= hi) if mode == 0 else self.gam > lo
        k, j = np.nonzero(m.T)
        return j.astype(np.int32), k.astype(np.int32), self.gam[j, k], int(m.sum())


def test_bfdr_selection_over_two_slabs_matches_global_assign_bFDR():
    """Host logic of summarise.select_bFDR_device when the traits are split over two contexts: counts and sums add up
    over slabs, the boundary ties are cut in global column-major order (lower slab first)."""
    from atlasqtl_b200 import summarise
    rng = np.random.default_rng(5)
    p, q = 40, 30
    gam = rng.uniform(size=(p, q)) ** 5
    gam[rng.uniform(size=(p, q)) < 0.05] = 0.97
    gam[3, 2] = gam[7, 20] = gam[9, 25] = gam[1, 11] = 0.9   # ties across the two slabs
    halves = [gam[:, :15], gam[:, 15:]]
    ctxs = [_PpiContext(h) for h in halves]
    for thres in (0.04, 0.05, 0.2):
        want = summarise.assign_bFDR(gam) < thres
        got = np.zeros_like(want)
        total = 0
        for rank in (0, 1):
            # a faithful sequential emulation: run rank r's algorithm with a communicator that adds the peer's values,
            # obtained by evaluating the same probes on the peer context
            peer = ctxs[1 - rank]

            class SeqComm:
                world_size = 2

                def __init__(self):
                    self.rank = rank
                    self.last_t = None

                def allreduce_sum(self, x):
                    x = np.asarray(x, dtype=np.float64)
                    if x.shape == (2,) and self.rank_probe is not None:
                        c, s = peer.ppi_count_sum(self.rank_probe)
                        return x + np.array([c, s])
                    if x.shape == (2,):   # tie counts: own slot filled, peer's slot from the peer's tie count
                        out = x.copy()
                        out[1 - rank] = self.peer_ties
                        return out
                    return x

                def allreduce_min(self, x):
                    return np.minimum(x, peer.ppi_next_above(self.t1))

            comm = SeqComm()
            ctx = ctxs[rank]
            # wrap the context so that the communicator knows which probe is in flight
            class Probe:
                p, q = ctx.p, ctx.q

                def ppi_count_sum(self, t):
                    comm.rank_probe = t
                    return ctx.ppi_count_sum(t)

                def ppi_next_above(self, t):
                    comm.t1 = t
                    return ctx.ppi_next_above(t)

                def ppi_collect(self, mode, lo, hi, capacity):
                    if mode == 0 and lo >= 0:   # the tie collection
                        comm.rank_probe = None
                        comm.peer_ties = peer.ppi_collect(0, lo, hi, capacity)[3]
                    return ctx.ppi_collect(mode, lo, hi, capacity)

            rows, cols, nsel = summarise.select_bFDR_device(Probe(), thres, comm=comm, k_first=15 * rank, p=p)
            got[rows, cols] = True
            total = nsel
        assert total == want.sum(), thres
        assert np.array_equal(got, want), thres
