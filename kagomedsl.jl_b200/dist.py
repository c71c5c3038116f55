"""Multi-GPU plumbing: walkers shard across ranks (one process per GPU, torch.distributed over
NCCL); the only exchange of the path is the sum of the per-bin observable accumulators."""
from __future__ import annotations

import os

import numpy as np


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)"""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_walkers(n_walkers_total: int, rank: int, world_size: int):
    """contiguous block of the global walker range owned by `rank`: (first_walker, count)"""
    base, rem = divmod(int(n_walkers_total), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the C ABI (kdsl_comm_unique_id): 128 opaque bytes for rank 0 to hand out"""
    import ctypes as C
    from . import _lib
    buf = (C.c_uint8 * _lib.COMM_ID_BYTES)()
    _lib.check(_lib.lib().kdsl_comm_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf)


def init_comm(engine) -> None:
    """Give `engine` its NCCL communicator inside libkdsl (one process per GPU): rank 0 draws the unique id, the
    already-initialised torch.distributed group (any backend) carries its 128 bytes to the other ranks, then every rank
    joins with kdsl_comm_init_rank.  After this `Engine.accumulators_allreduce` / `accumulators(mc)` reduce through the
    C ABI and torch.distributed is no longer on the path.  No-op for a single rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    rank, n = dist.get_rank(), dist.get_world_size()
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    engine.comm_init_rank(n, rank, box[0])


class Group:
    """Single-process multi-GPU job (the layout a Julia host uses): one Engine per device, one NCCL communicator made
    by kdsl_comm_init_all, accumulators summed by kdsl_group_accumulators_allreduce."""

    def __init__(self, engines):
        import ctypes as C
        from . import _lib
        self.engines = list(engines)
        self._arr = (C.c_void_p * len(self.engines))(*[e._h for e in self.engines])
        _lib.check(_lib.lib().kdsl_comm_init_all(len(self.engines), self._arr))

    def sweep(self, n_sweeps: int, thermalization: int = -1) -> None:
        for e in self.engines:                           # asynchronous launches: the GPUs run concurrently
            e.sweep(n_sweeps, thermalization)

    def accumulators(self) -> np.ndarray:
        import ctypes as C
        from . import _lib
        out = np.zeros(_lib.N_ACC)
        _lib.check(_lib.lib().kdsl_group_accumulators_allreduce(len(self.engines), self._arr, out.ctypes.data_as(C.c_void_p)))
        return out


def allreduce_sum(vec: np.ndarray, device: int = 0) -> np.ndarray:
    """sum a small float64 vector over all ranks (NCCL on GPU tensors, gloo on CPU tensors);
    identity when torch.distributed is not initialised"""
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return vec
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return vec
    import torch
    if dist.get_backend() == "nccl":
        t = torch.as_tensor(np.asarray(vec, dtype=np.float64), device=f"cuda:{device}")
    else:
        t = torch.as_tensor(np.asarray(vec, dtype=np.float64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
