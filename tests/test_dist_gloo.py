"""CPU, world_size = 2 over gloo: the trait-slab sharding and the all-reduce protocol of atlasqtl_b200.core /
dist reproduce the single-process run (the CUDA context is replaced by the oracle-backed test double)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from problems import make_problem

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, tmpdir, anneal):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    from atlasqtl_b200 import core
    from atlasqtl_b200.dist import TorchComm, slab_bounds
    from fake_context import OracleSweepContext
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    X, Y, hyper, init = make_problem(100, 75, 21, p_act=10, q_act=20, maf=0.2, p0=(5, 25))
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


@pytest.mark.parametrize("anneal", [(1, 2, 10), None])
def test_two_slabs_reproduce_single_process(oracle_built, tmp_path, anneal):
    from atlasqtl_b200 import core
    from fake_context import OracleSweepContext
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path), anneal), nprocs=2, join=True)
    d = np.load(tmp_path / "dist.npz")
    X, Y, hyper, init = make_problem(100, 75, 21, p_act=10, q_act=20, maf=0.2, p0=(5, 25))
    tr = []
    one = core.atlasqtl_global_local_core_(Y, X, Y.shape[1], anneal, 1, 0.1, 1000, 0, hyper, init, debug=True, trace=tr,
                                           context_factory=lambda X_, Y_: OracleSweepContext(X_, Y_))
    assert int(d["it"]) == one["it"]
    lbs = np.array([r["lb"] for r in tr if r["lb"] is not None])
    np.testing.assert_allclose(d["lbs"], lbs, rtol=1e-11)
    assert np.abs(d["gam"] - one["gam_vb"]).max() <= 1e-10
    np.testing.assert_allclose(d["theta"], one["theta_vb"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(d["zeta"], one["zeta_vb"], rtol=1e-9, atol=1e-11)
