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
