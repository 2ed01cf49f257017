"""Trait-sharded multi-GPU execution: one process per GPU under torch.distributed (NCCL over NVLink).

Inside a sweep, column k of every array is touched by trait k only (reference src/coreLoop.cpp:58-85), so
the q traits are split into contiguous slabs, one per rank; X is replicated.  The only exchange is ONE
all-reduce per sweep of [rowSums of the Z part (p) | sum(gam_vb) | sum_k tau_k colSums(m2_beta)_k] and one
scalar-sized all-reduce of ELBO partial sums on ELBO iterations.
"""
import numpy as np


def slab_bounds(q, rank, world_size):
    """Contiguous, nearly equal slabs of traits: [k_first, k_last)."""
    base, rem = divmod(q, world_size)
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


class _DeviceArray:
    """A raw device pointer dressed as a CUDA array (float64, 1-D) so that torch can alias it without a copy."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class TorchComm:
    """allreduce_sum over the default process group.  NCCL needs device tensors: the (small) buffer goes
    host -> device -> all-reduce -> host; `gloo` (CPU tests) reduces the host tensor directly."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._torch, self._dist = torch, dist
        self.rank, self.world_size = dist.get_rank(), dist.get_world_size()
        self.backend = dist.get_backend()
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu"))
        if self.device.type != "cuda":
            self.allreduce_sum_device = None   # (the core checks for a callable: host buffers go through allreduce_sum)

    def allreduce_sum(self, x):
        t = self._torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        if self.device.type == "cuda":
            t = t.to(self.device, non_blocking=False)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def allreduce_sum_device(self, dev_ptr, count):
        """In-place NCCL all-reduce of `count` doubles at device address `dev_ptr` (a buffer owned by the CUDA library,
        e.g. aq_rowsums_zpart_dev), then one download.  Only with the nccl backend (AttributeError otherwise, which the
        core treats as "not available")."""
        if self.device.type != "cuda":
            raise AttributeError("allreduce_sum_device needs the nccl backend")
        t = self._torch.as_tensor(_DeviceArray(dev_ptr, count), device=self.device)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def allreduce_min(self, x):
        t = self._torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        if self.device.type == "cuda":
            t = t.to(self.device, non_blocking=False)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MIN)
        return t.cpu().numpy()

    def gather_result(self, res, q):
        """Assemble the full-width output object on every rank (all_gather of the p x q_local slabs)."""
        out = dict(res)
        for key in ("gam_vb", "beta_vb", "mu_beta_vb"):
            if res.get(key) is not None:
                parts = [None] * self.world_size
                self._dist.all_gather_object(parts, res[key])
                out[key] = np.asfortranarray(np.concatenate(parts, axis=1))
        for key in ("zeta_vb", "tau_vb", "sig2_beta_vb", "eta_vb", "kappa_vb"):
            if res.get(key) is not None:
                parts = [None] * self.world_size
                self._dist.all_gather_object(parts, res[key])
                out[key] = np.concatenate(parts)
        out["q"] = q
        return out
