"""Multi-GPU plan for the sampling path: time-range sharding of the loader-batch stream.

In the stateless form a batch's sampled neighbourhood depends only on the immutable store and
the batch's position in the stream, so the batch stream partitions into contiguous ranges with
no exchange step: rank r of W owns loader batches [nb*r//W, nb*(r+1)//W); the time-sorted store
and the per-node adjacency are replicated per GPU (1.6 GB + features at 1e8 edges, against
180 GB of HBM).  Sampling has NO data-path collective; torch.distributed is used only for the
barrier / max-over-ranks timing and for gathering per-shard summaries.
The one real exchange of the path is the TGN node-memory join at shard boundaries
(`join_node_memory`): the rows each shard touched, packed and all-gathered once.
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


def _pack_rows(memory: torch.Tensor, last_update: torch.Tensor, ids: torch.Tensor,
               cap: int) -> torch.Tensor:
    """uint8 [cap, 16 + 4 M]: packed records of the rows `ids` (tgm_join_pack); padding ids -1."""
    M = memory.shape[1]
    rb = 16 + 4 * M
    n = ids.numel()
    if memory.is_cuda:
        from tgm_b200 import _cabi
        rows = torch.empty((cap, rb), dtype=torch.uint8, device=memory.device)
        _cabi.check(_cabi.lib.tgm_join_pack(
            memory.data_ptr(), last_update.data_ptr(), M, ids.data_ptr() if n else None, n, cap,
            rows.data_ptr(), _cabi.current_stream(memory.device)))
        return rows
    # host tensors: the world-size-2 gloo tests of the exchange protocol (no GPU in that container)
    rows = torch.zeros((cap, rb), dtype=torch.uint8)
    rows[:, :4] = torch.full((cap,), -1, dtype=torch.int32).view(torch.uint8).view(cap, 4)
    if n:
        idl = ids.long()
        rows[:n, :4] = ids.to(torch.int32).contiguous().view(torch.uint8).view(n, 4)
        rows[:n, 8:16] = last_update[idl].contiguous().view(torch.uint8).view(n, 8)
        rows[:n, 16:] = memory[idl].contiguous().view(torch.uint8).view(n, 4 * M)
    return rows


def _scatter_rows(rows: torch.Tensor, n: int, memory: torch.Tensor,
                  last_update: torch.Tensor) -> None:
    M = memory.shape[1]
    if not n:
        return
    if memory.is_cuda:
        from tgm_b200 import _cabi
        _cabi.check(_cabi.lib.tgm_join_scatter(
            rows.data_ptr(), n, M, memory.shape[0], memory.data_ptr(), last_update.data_ptr(),
            _cabi.current_stream(memory.device)))
        return
    r = rows[:n].contiguous()
    ids = r[:, :4].contiguous().view(torch.int32).view(n).long()
    last_update[ids] = r[:, 8:16].contiguous().view(torch.int64).view(n)
    memory[ids] = r[:, 16:].contiguous().view(torch.float32).view(n, M)


def join_node_memory(memory: torch.Tensor, last_update: torch.Tensor,
                     touched_ids: torch.Tensor) -> dict:
    """Reconcile TGN node memory at a shard join, in place: the same result as
    `merge_node_memory`, moving only the rows that changed.

    `touched_ids`: the UNIQUE node ids (int32) this rank's shard wrote since the common snapshot
    (endpoints of its events: tgm/nn/encoder/tgn.py:192-216 writes memory[n_id] and
    last_update[n_id] for exactly those).  Protocol: (1) all-gather of the per-rank counts (one
    small message, the only host sync); (2) `tgm_join_pack` gathers this rank's rows into records
    {id, last_update, memory row}, padded to the largest count; (3) ONE all-gather of the record
    blocks over NCCL/NVLink; (4) `tgm_join_scatter` applies the blocks in ascending rank order, so
    the later time shard wins a row two shards touched and untouched rows keep the snapshot.
    Bytes on the wire per rank: max_count * (16 + 4 M) instead of three dense [N]-row reductions.
    Returns the counts and byte figures of this join."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {'counts': [int(touched_ids.numel())], 'row_bytes': 16 + 4 * memory.shape[1],
                'recv_bytes': 0}
    world = dist.get_world_size()
    n = int(touched_ids.numel())
    cnt = torch.tensor([n], dtype=torch.int64, device=memory.device)
    counts = torch.empty(world, dtype=torch.int64, device=memory.device)
    dist.all_gather_into_tensor(counts, cnt)
    counts = [int(c) for c in counts.tolist()]
    cap = max(counts)
    rb = 16 + 4 * memory.shape[1]
    if cap == 0:
        return {'counts': counts, 'row_bytes': rb, 'recv_bytes': 0}
    mine = _pack_rows(memory, last_update, touched_ids.to(torch.int32).contiguous(), cap)
    blocks = torch.empty((world, cap, rb), dtype=torch.uint8, device=memory.device)
    dist.all_gather_into_tensor(blocks.view(-1), mine.view(-1))
    for r in range(world):
        _scatter_rows(blocks[r], counts[r], memory, last_update)
    return {'counts': counts, 'row_bytes': rb, 'recv_bytes': sum(counts) * rb}


def bench_memory_join(num_nodes: int, memory_dim: int, src: torch.Tensor, dst: torch.Tensor,
                      join_edges: int, device, reps: int = 5) -> dict:
    """Time the shard-join exchange (bench.py `collective`): each rank's touched set = endpoints
    of the first `join_edges` stream edges of its shard (device dedup: bitmap + popcount scan),
    then `join_node_memory`; the dense three-all-reduce form is timed beside it.  CUDA events,
    max over ranks; checks that both forms leave identical memory."""
    from tgm_b200.hooks.dedup import _BatchIdSet
    world, rank = dist.get_world_size(), dist.get_rank()
    g = torch.Generator(device=device).manual_seed(1234)  # the common snapshot, same on all ranks
    snap = torch.randn((num_nodes, memory_dim), generator=g, device=device)
    snap_lu = torch.zeros(num_nodes, dtype=torch.int64, device=device)
    s, d = src[:join_edges].contiguous(), dst[:join_edges].contiguous()

    def shard_state():
        """What this rank's shard would leave: its touched rows rewritten (rank-specific)."""
        mem, lu = snap.clone(), snap_lu.clone()
        ids = _BatchIdSet(num_nodes, device).unique([(s, False), (d, False)])
        mem[ids.long()] = mem[ids.long()] * 0.5 + (rank + 1)
        lu[ids.long()] = 1000 * (rank + 1)
        return mem, lu

    def timed(fn):
        best = []
        for _ in range(reps):
            mem, lu = shard_state()
            dist.barrier()
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            info = fn(mem, lu)
            e1.record()
            torch.cuda.synchronize(device)
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best.append(float(t.item()))
        return min(best), sorted(best)[len(best) // 2], info, mem, lu

    def sparse(mem, lu):
        ids = _BatchIdSet(num_nodes, device).unique([(s, False), (d, False)])
        return join_node_memory(mem, lu, ids)

    def dense(mem, lu):
        touched = torch.zeros(num_nodes, dtype=torch.bool, device=device)
        touched[s.long()] = True
        touched[d.long()] = True
        merge_node_memory(mem, lu, touched)
        return None

    t_best, t_med, info, mem_a, lu_a = timed(sparse)
    d_best, d_med, _, mem_b, lu_b = timed(dense)
    same = bool(torch.equal(mem_a, mem_b) and torch.equal(lu_a, lu_b))
    same_t = torch.tensor([int(same)], device=device)
    dist.all_reduce(same_t, op=dist.ReduceOp.MIN)
    recv = info['recv_bytes']
    return {
        'what': 'TGN node-memory join at a shard boundary: device dedup of the touched endpoints, '
                'pack, ONE all-gather of the touched rows over NCCL, scatter in rank order',
        'nodes': num_nodes, 'memory_dim': memory_dim, 'edges_per_shard': int(s.numel()),
        'touched_rows_per_rank': info['counts'], 'row_bytes': info['row_bytes'],
        'ms': t_best, 'ms_median': t_med,
        'algorithmic_bytes_per_gpu': recv,
        'algorithmic_gbs_per_gpu': recv / (t_best * 1e-3) / 1e9,
        'dense_allreduce_ms': d_best, 'dense_allreduce_ms_median': d_med,
        'equals_dense_form': bool(same_t.item()), 'world': world,
    }


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
