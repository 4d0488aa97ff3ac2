"""Multi-GPU plumbing for batch-sharded ensembles.

One process per GPU.  Independent trajectories shard by batch with no data-path collective; an
adaptive ensemble that shares ONE dt (the reference's semantics when ``u`` is ``(B, n)`` and
``lin_op`` is ``(n,)``: global max and global 2-norms over the whole batch, solveras.py:451-454)
needs the three error-norm scalars combined across ranks once per trial:

    red[0] = max |u+|^2          -> all_reduce(MAX)   (before the masked sums: the mask needs it)
    red[1] = sum_mask |u+|^2     -> all_reduce(SUM)
    red[2] = sum_mask |err|^2    -> all_reduce(SUM)

The functions here are backend agnostic (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of the batch owned by `rank` (ragged batches allowed)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(u: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """This rank's rows of a (B, ...) ensemble."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(u.shape[0], rank, world)
    return u[lo:hi]


def allreduce_error_scalars(red: torch.Tensor, group=None, between: Optional[Callable[[], None]] = None) -> None:
    """Combine the per-rank reduction scalars in place.

    `red` is the 3-element float64 view of the control block; `between` runs after the MAX
    exchange and before the SUM exchange (on the GPU: the masked-sum kernel, which needs the global max).
    """
    if red.numel() != 3 or red.dtype != torch.float64:
        raise ValueError("red must be 3 float64 values")
    dist.all_reduce(red[0:1], op=dist.ReduceOp.MAX, group=group)
    if between is not None:
        between()
    dist.all_reduce(red[1:3], op=dist.ReduceOp.SUM, group=group)
