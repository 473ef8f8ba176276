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


class Rank0Gather:
    """Result gather of the utterance-sharded run: every rank's padded waveforms + lengths land on rank 0, asynchronously.

    bench.py's steps are independent micro-batches, so the gather of step i overlaps the compute of step i+1: the collective is issued
    with async_op=True (it runs on the process group's own NCCL stream, ordered after the producing kernels of the current stream) from
    one of two send buffers, and the host only waits for a buffer's previous gather before reusing it.  No per-step host sync, no
    per-step allocation, and only rank 0 receives (NVLink traffic world-1 payloads per step instead of world*(world-1) for an
    all-gather).  The time capacity is negotiated once (one all_reduce(MAX) + .item() at the first submit, with head-room); a later step
    that outgrows it raises (the caller then calls reset() on EVERY rank, a collective decision).  Works on any backend (gloo in the
    CPU tests; with NCCL, work.wait() orders the current stream after the collective without blocking the host)."""

    def __init__(self, group=None, headroom: float = 1.10, align: int = 256):
        self.group, self.headroom, self.align = group, headroom, align
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.cap = 0
        self.send = [None, None]
        self.send_len = [None, None]
        self.recv = None
        self.recv_len = None
        self.work = [None, None]
        self.turn = 0

    def _negotiate(self, local: torch.Tensor):
        need = torch.tensor([local.shape[1]], dtype=torch.int64, device=local.device)
        dist.all_reduce(need, op=dist.ReduceOp.MAX, group=self.group)
        cap = int(int(need.item()) * self.headroom)
        self.cap = (cap + self.align - 1) // self.align * self.align
        shape = (local.shape[0], self.cap) + tuple(local.shape[2:])
        self.send = [local.new_zeros(shape) for _ in range(2)]
        self.send_len = [torch.zeros(local.shape[0], dtype=torch.int64, device=local.device) for _ in range(2)]
        if self.rank == 0:
            self.recv = [[local.new_empty(shape) for _ in range(self.world)] for _ in range(2)]
            self.recv_len = [[torch.empty(local.shape[0], dtype=torch.int64, device=local.device) for _ in range(self.world)] for _ in range(2)]

    def submit(self, local: torch.Tensor, lengths: torch.Tensor):
        """local [b, T_r, ...] (T_r may differ per rank and per step), lengths int64 [b]; returns immediately."""
        if self.world == 1:
            return
        if self.cap == 0:
            self._negotiate(local)
        if local.shape[1] > self.cap:
            raise RuntimeError(f"Rank0Gather: {local.shape[1]} time steps exceed the negotiated capacity {self.cap}; call reset() on every rank")
        t = self.turn
        if self.work[t] is not None:
            for w in self.work[t]:
                w.wait()
        n = local.shape[1]
        self.send[t][:, :n].copy_(local)
        self.send_len[t].copy_(lengths)
        kw = dict(dst=0, group=self.group, async_op=True)
        self.work[t] = [dist.gather(self.send_len[t], self.recv_len[t] if self.rank == 0 else None, **kw),
                        dist.gather(self.send[t], self.recv[t] if self.rank == 0 else None, **kw)]
        self.turn ^= 1

    def reset(self):
        """Collective: drain and renegotiate the capacity at the next submit."""
        self.flush()
        self.cap = 0

    def flush(self):
        for t in (0, 1):
            if self.work[t] is not None:
                for w in self.work[t]:
                    w.wait()
                self.work[t] = None

    def last(self):
        """Rank 0, after flush(): (payload [world*b, cap, ...], lengths [world*b]) of the most recent submit."""
        t = self.turn ^ 1
        if self.rank != 0 or self.recv is None:
            return None
        return torch.cat(self.recv[t], dim=0), torch.cat(self.recv_len[t], dim=0)
