#!/usr/bin/env python
"""BASELINE config 4: TGN node memory on a time-sharded CTDG with the memory join at the shard
boundary.  Launch with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
        --master-port 29533 bench_tgn_shard.py --batches 500

Every rank owns a contiguous time range of the stream (tgm_b200/parallel.py), samples k=10 recent
neighbours for its batches from the replicated store (one pre-sampled window), runs
TGNMemory.forward + update_state per batch, then all ranks reconcile memory with
`merge_node_memory` (the single collective of the path).  Prints one JSON line on rank 0; times
are CUDA-event times, max over ranks."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tgm_b200 import RecencyCSR  # noqa: E402
from tgm_b200.core.storage import DeviceCOOStorage  # noqa: E402
from tgm_b200.nn import TGNMemory  # noqa: E402
from tgm_b200.parallel import max_over_ranks, merge_node_memory, shard_batches, sum_over_ranks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--edges', type=int, default=20_000_000)
    ap.add_argument('--nodes', type=int, default=1_000_000)
    ap.add_argument('--dim', type=int, default=16)
    ap.add_argument('--k', type=int, default=10)
    ap.add_argument('--batch-size', type=int, default=200)
    ap.add_argument('--batches', type=int, default=500, help='loader batches per rank')
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=dev)
    E, N, D, bs, k = a.edges, a.nodes, a.dim, a.batch_size, a.k
    g = torch.Generator(device=dev).manual_seed(0)
    src = torch.randint(0, N, (E,), generator=g, device=dev, dtype=torch.int32)
    dst = torch.randint(0, N, (E,), generator=g, device=dev, dtype=torch.int32)
    t = torch.arange(E, device=dev, dtype=torch.int64) * 3  # unique times: TGN parity domain
    x = torch.randn(E, D, generator=g, device=dev)
    store = DeviceCOOStorage.from_device_tensors(src, dst, t, x, N)
    csr = RecencyCSR(store, bs, colocate_x=True)
    shard = shard_batches(E, bs, rank, world)
    lo = shard.edge_lo
    hi = min(shard.edge_hi, lo + a.batches * bs)
    torch.manual_seed(0)
    mem = TGNMemory(N, D, 100, 100).to(dev)
    mem.train()
    mem.reset_state()
    touched = torch.zeros(N, dtype=torch.bool, device=dev)

    def run_shard():
        hops = csr.sample_window(lo, hi, [k])  # neighbourhoods of every batch of the shard
        for b_lo in range(lo, hi, bs):
            b_hi = min(b_lo + bs, hi)
            s, d = src[b_lo:b_hi], dst[b_lo:b_hi]
            n_id = torch.cat([s, d]).long()
            mem(n_id)
            mem.update_state(s, d, t[b_lo:b_hi], x[b_lo:b_hi])
            touched[n_id] = True
        return hops

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    run_shard()  # warm-up (allocations, cuBLAS heuristics)
    for _ in range(2):  # warm-up of the collective (communicator setup, NVLS buffers)
        merge_node_memory(mem.memory.clone(), mem.last_update.clone(), touched)
    mem.reset_state()
    touched.zero_()
    barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    run_shard()
    e[1].record()
    memory, last_update = mem.memory, mem.last_update
    merge_node_memory(memory, last_update, touched)
    e[2].record()
    barrier()
    shard_ms = max_over_ranks(e[0].elapsed_time(e[1]), dev)
    merge_ms = max_over_ranks(e[1].elapsed_time(e[2]), dev)
    events = sum_over_ranks(float(hi - lo), dev)
    check = sum_over_ranks(float(memory.double().abs().sum().item()), dev) / world
    if rank == 0:
        row_bytes = N * (100 * 4 + 8 + 4)
        print(json.dumps({
            'row': 'config 4: TGN memory, time-sharded, memory join at the shard boundary',
            'n_gpus': world, 'batches_per_rank': (hi - lo) // bs, 'nodes': N,
            'events_per_s': events / ((shard_ms + merge_ms) * 1e-3),
            'shard_ms': shard_ms, 'merge_ms': merge_ms,
            'merge_algorithmic_GB_per_s_per_gpu': row_bytes / (merge_ms * 1e-3) / 1e9,
            'merged_memory_l1': check,
            'note': 'merge = MAX all-reduce int32[N] + SUM all-reduce f32[N,100] + MAX all-reduce '
                    'int64[N] over NCCL; per-batch memory work is launch-bound at bs=200'}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
