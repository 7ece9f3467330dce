"""Host-side plumbing for the data-parallel configuration: states shard over ranks with
no data-path collective; the only exchange is one SUM all-reduce of the contiguous theta-gradient
buffer in the critic regression step.  Pure torch.distributed (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of n units (states) for `rank`: the first n % world ranks
    get one extra unit.  Every unit belongs to exactly one rank."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_grad_(grad: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """In-place SUM all-reduce of a per-rank gradient that was already scaled by 1/B_total
    (rlc_critic_grads' B_total argument), which reproduces the reference's batch *mean*
    (forwardkl_network.py:140) without a second pass."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    return grad


def global_mean_from_shards(per_state: torch.Tensor, n_total: int,
                            group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """mean over ALL states of a per-state quantity held shard-wise (e.g. the FKL/RKL policy loss,
    forwardkl_network.py:194): sum locally, SUM all-reduce one scalar, divide by the global count."""
    tot = per_state.sum().reshape(1).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    return tot / float(n_total)
