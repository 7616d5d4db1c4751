"""Multi-GPU plan for the sampling path: time-range sharding of the loader-batch stream.

In the stateless form a batch's sampled neighbourhood depends only on the immutable store and
the batch's position in the stream, so the batch stream partitions into contiguous ranges with
no exchange step: rank r of W owns loader batches [nb*r//W, nb*(r+1)//W); the time-sorted store
and the per-node adjacency are replicated per GPU (1.6 GB + features at 1e8 edges, against
180 GB of HBM).  There is NO data-path collective; torch.distributed is used only for the
barrier / max-over-ranks timing and for gathering per-shard summaries.
The reference has no multi-device path at all (single process, `tgm/data/loader.py:64-184`).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List, Tuple

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class TimeRangeShard:
    """Contiguous range of loader batches (and the stream edges they cover) owned by one rank."""
    rank: int
    world: int
    batch_lo: int
    batch_hi: int
    edge_lo: int
    edge_hi: int

    @property
    def num_batches(self) -> int:
        return self.batch_hi - self.batch_lo

    @property
    def num_edges(self) -> int:
        return self.edge_hi - self.edge_lo

    def windows(self, window_batches: int, batch_size: int) -> Iterator[Tuple[int, int]]:
        """[e_lo, e_hi) edge windows of at most `window_batches` loader batches, in stream order;
        every window starts on a batch boundary (what tgm_csr_sample_edges requires)."""
        step = window_batches * batch_size
        for lo in range(self.edge_lo, self.edge_hi, step):
            yield lo, min(lo + step, self.edge_hi)


def shard_batches(num_edges: int, batch_size: int, rank: int, world: int,
                  e_start: int = 0) -> TimeRangeShard:
    """The shard of rank `rank`: batches are [e_start + i*bs, e_start + (i+1)*bs) as the loader
    cuts them (tgm/data/loader.py:136-148); shards differ in size by at most one batch."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside [0, {world})')
    if batch_size <= 0 or num_edges < 0:
        raise ValueError('batch_size must be > 0 and num_edges >= 0')
    n = max(0, num_edges - e_start)
    nb = (n + batch_size - 1) // batch_size
    b_lo, b_hi = nb * rank // world, nb * (rank + 1) // world
    e_lo = e_start + b_lo * batch_size
    e_hi = min(e_start + b_hi * batch_size, num_edges)
    return TimeRangeShard(rank, world, b_lo, b_hi, e_lo, max(e_lo, e_hi))


def current_shard(num_edges: int, batch_size: int, e_start: int = 0) -> TimeRangeShard:
    """Shard of this process under torch.distributed (rank 0 of 1 when not initialised)."""
    if dist.is_available() and dist.is_initialized():
        return shard_batches(num_edges, batch_size, dist.get_rank(), dist.get_world_size(), e_start)
    return shard_batches(num_edges, batch_size, 0, 1, e_start)


def gather_shard_summaries(values: List[float], device='cpu') -> torch.Tensor:
    """All-gather a small per-rank vector (counts, checksums, timings) -> [world, len(values)].
    Works on gloo (CPU tests) and nccl alike; not on the data path."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()):
        return t[None]
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out)


def max_over_ranks(value: float, device='cpu') -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device='cpu') -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def merge_node_memory(memory: torch.Tensor, last_update: torch.Tensor,
                      touched: torch.Tensor) -> None:
    """Reconcile TGN-style node memory at a shard join (BASELINE config 4), in place.

    Every rank advanced ITS time-range shard of the event stream starting from the same memory
    snapshot and touched a subset of the nodes (`touched` bool[N]: endpoints of its shard's
    events).  Shards are ordered in time by rank, so for each node the row written by the highest
    rank that touched it is the most recent one; nodes nobody touched keep the common value.

    This is the one collective of the hot path: a rank-id MAX all-reduce over int32[N] elects the
    owner of every row, then one SUM all-reduce over the owner-masked memory [N, M] (and one MAX
    over last_update) delivers the rows -- the all-gather of the touched rows, expressed as
    reductions so NVSwitch can combine in-network (NVLS).  Time-sharded training of a memory model
    is an approximation of the sequential reference (tgm/nn/encoder/tgn.py processes the stream
    strictly in order): inside a shard a node sees only its own shard's updates until the join.
    Works on NCCL (device tensors) and gloo (CPU tests) alike.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    rank = dist.get_rank()
    owner = torch.where(touched, rank + 1, 0).to(torch.int32)
    dist.all_reduce(owner, op=dist.ReduceOp.MAX)
    mine = owner == rank + 1
    nobody = owner == 0
    keep = mine | (nobody if rank == 0 else torch.zeros_like(nobody))  # rank 0 speaks for untouched rows
    contrib = torch.where(keep[:, None], memory, torch.zeros((), dtype=memory.dtype, device=memory.device))
    dist.all_reduce(contrib, op=dist.ReduceOp.SUM)
    lu = torch.where(keep, last_update, torch.zeros((), dtype=last_update.dtype, device=last_update.device))
    dist.all_reduce(lu, op=dist.ReduceOp.MAX)
    memory.copy_(contrib)
    last_update.copy_(lu)


def average_gradients(params) -> None:
    """Data-parallel training over time shards: average `.grad` of the given parameters across
    ranks with ONE all-reduce over a flat buffer (a training step's gradients are a few MB: one
    launch-latency-sized message, NVLS-reducible on NVSwitch).  Parameters without a gradient on
    this rank contribute zeros; afterwards every rank holds the same gradients.  No-op outside
    torch.distributed.  Works on NCCL and on gloo (CPU tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
                      for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    for p, v in zip(params, flat.split([p.numel() for p in params])):
        if p.grad is None:
            p.grad = v.view_as(p).clone()
        else:
            p.grad.copy_(v.view_as(p))
