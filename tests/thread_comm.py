"""Test double for dist.TorchComm: `world` ranks as THREADS of one process (one CUDA context each on the same GPU), with
a barrier-based all-reduce that sums the contributions in rank order -- the protocol of the trait-slab sharding without
NCCL, so that the hardware path of every slab (real CUDA contexts) can be checked on a single-GPU box."""
import threading

import numpy as np


class ThreadGroup:
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world

    def comm(self, rank):
        return ThreadComm(self, rank)


class ThreadComm:
    def __init__(self, group, rank):
        self.g, self.rank, self.world_size = group, rank, group.world

    def _reduce(self, x, op):
        self.g.slots[self.rank] = np.array(x, dtype=np.float64)
        self.g.barrier.wait()
        out = self.g.slots[0].copy()
        for r in range(1, self.world_size):
            out = op(out, self.g.slots[r])
        self.g.barrier.wait()
        return out

    def allreduce_sum(self, x):
        return self._reduce(x, np.add)

    def allreduce_min(self, x):
        return self._reduce(x, np.minimum)
