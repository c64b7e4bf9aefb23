"""Multi-GPU plumbing: scenes are independent, so they shard across ranks with no
data-path collective.  torch.distributed (NCCL on GPUs, gloo in CPU tests)
is used only for the barrier and for aggregating timings / counters."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_info():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n_total: int, rank: int, world: int):
    """contiguous scene range [lo, hi) of `rank` (sizes differ by at most one)"""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_timing(ms: float, units: int, device=None):
    """(max over ranks of ms, sum over ranks of units); identity when not distributed"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms), int(units)
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    n = torch.tensor([units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), int(n.item())


def barrier(device=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device is not None and torch.cuda.is_available():
        torch.cuda.synchronize(device)
