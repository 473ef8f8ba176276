"""Utterance sharding over the GPUs of one box (one process per GPU, torch.distributed over NCCL / NVLink).

The path has no exchange step: utterances are independent in eval mode (BatchNorm uses running statistics), so each
rank runs whole micro-batches and the only collective is the result gather.  The unit of work
is a fixed micro-batch, not an utterance, because the reference's outputs depend on batch composition (padding leaks
through its unmasked convolutions): every GPU count must process the same micro-batches.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_microbatches(n_micro: int, rank: int, world: int) -> List[int]:
    """Contiguous block assignment of micro-batch indices to ranks (remainder to the lowest ranks)."""
    base, rem = divmod(n_micro, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def gather_padded(local: torch.Tensor, lengths: torch.Tensor, group=None):
    """All-gather a per-rank [b, T_r, ...] result whose time extent differs across ranks.

    Lengths first (int64 [b]), then the payload padded to the global max T.  Returns (gathered [world*b, Tmax, ...],
    lengths [world*b]) on every rank, rows ordered by rank.  Requires the same b on every rank."""
    world = dist.get_world_size(group)
    if world == 1:
        return local, lengths
    tmax = torch.tensor([local.shape[1]], dtype=torch.int64, device=local.device)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX, group=group)
    T = int(tmax.item())
    if local.shape[1] < T:
        pad = local.new_zeros((local.shape[0], T - local.shape[1]) + tuple(local.shape[2:]))
        local = torch.cat([local, pad], dim=1)
    local = local.contiguous()
    out = local.new_empty((world * local.shape[0],) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local, group=group)
    lens = lengths.new_empty(world * lengths.shape[0])
    dist.all_gather_into_tensor(lens, lengths.contiguous(), group=group)
    return out, lens
