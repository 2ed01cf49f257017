"""Worker of tests/test_gpu_multi_slab.py::test_nccl_two_ranks_reproduce_single_gpu (one process per GPU under torchrun)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import torch
    import torch.distributed as dist
    from atlasqtl_b200 import api, core
    from atlasqtl_b200.dist import TorchComm, slab_bounds
    from problems import make_problem
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = TorchComm()
    X, Y, hyper, init = make_problem(300, 240, 2500, seed=7)
    q = Y.shape[1]
    k0, k1 = slab_bounds(q, rank, world)
    tr = []
    out = core.atlasqtl_global_local_core_(np.asfortranarray(Y[:, k0:k1]), X, q, (1, 2, 5), 1, 0.1, 40, 0, hyper, init,
                                           debug=True, comm=comm, slab=(k0, k1), device=local, trace=tr)
    full = comm.gather_result(out, q)
    # the public entry point with DEFAULT hyper / init and no seed: replicated state must agree across ranks
    res = api.atlasqtl(Y[:, :64], X[:, :80], (2, 10), anneal=(1, 2, 3), maxit=6, verbose=0, device=local, comm=comm)
    th = torch.tensor(res["theta_vb"], device="cuda")
    lo, hi = th.clone(), th.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    consistent = bool(torch.equal(lo, hi))
    if rank == 0:
        np.savez(sys.argv[1], gam=full["gam_vb"], theta=full["theta_vb"], it=out["it"],
                 lbs=np.array([r["lb"] for r in tr if r["lb"] is not None]), default_init_consistent=consistent)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
