#!/usr/bin/env python
"""BASELINE config 4: TGN node memory on a time-sharded CTDG with the memory join at the shard
boundary.  Launch with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
        --master-port 29533 bench_tgn_shard.py --batches 500

Every rank owns a contiguous time range of the stream (tgm_b200/parallel.py), samples k=10 recent
neighbours for its batches from the replicated store (one pre-sampled window), runs the model step
of examples/linkproppred/tgn.py per batch (device dedup -> TGNMemory.forward -> GraphAttentionEmbedding
-> update_state; `--no-embedding` keeps the memory state machine only), then all ranks reconcile memory with
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
from tgm_b200.hooks.dedup import _BatchIdSet, compact_frontier  # noqa: E402
from tgm_b200.nn import GraphAttentionEmbedding, TGNMemory  # noqa: E402
from tgm_b200.parallel import (average_gradients, max_over_ranks, merge_node_memory,  # noqa: E402
                               shard_batches, sum_over_ranks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--edges', type=int, default=20_000_000)
    ap.add_argument('--nodes', type=int, default=1_000_000)
    ap.add_argument('--dim', type=int, default=16)
    ap.add_argument('--k', type=int, default=10)
    ap.add_argument('--batch-size', type=int, default=200)
    ap.add_argument('--batches', type=int, default=500, help='loader batches per rank')
    ap.add_argument('--no-embedding', dest='embedding', action='store_false',
                    help='memory state machine only (the first measurement of the round)')
    ap.add_argument('--train', action='store_true',
                    help='training step per batch: link decoder + BCE loss (positives vs the batch\'s '
                         'rolled destinations), tgm_gae_backward + tgm_tgn_backward, gradient '
                         'all-reduce across the time shards, Adam')
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
    enc = GraphAttentionEmbedding(100, 100, D, mem.time_enc)
    enc.conv.dropout = 0.0  # the B200 path trains with dropout 0
    enc = enc.to(dev).train(a.train)
    mem.train()
    mem.reset_state()
    touched = torch.zeros(N, dtype=torch.bool, device=dev)
    decoder = torch.nn.Sequential(torch.nn.Linear(200, 100), torch.nn.ReLU(),
                                  torch.nn.Linear(100, 1)).to(dev)  # the example's LinkPredictor
    params = list({id(q): q for m_ in (mem, enc, decoder) for q in m_.parameters()}.values())
    opt = torch.optim.Adam(params, lr=1e-4)
    bce = torch.nn.functional.binary_cross_entropy_with_logits
    if a.train and not a.embedding:
        raise SystemExit('--train needs the embedding (drop --no-embedding)')

    def train_shard():
        """examples/linkproppred/tgn.py:70-121 per batch on this rank's shard; the gradients are
        averaged over the shards (data-parallel in time) before every Adam step."""
        hop = csr.sample_window(lo, hi, [k])[0]
        loss = None
        for b_lo in range(lo, hi, bs):
            b_hi = min(b_lo + bs, hi)
            s, d = src[b_lo:b_hi], dst[b_lo:b_hi]
            r0, r1 = 2 * (b_lo - lo), 2 * (b_hi - lo)
            nbr = hop.nbr_nids[r0:r1].reshape(-1)
            ids = _BatchIdSet(N, dev)
            uniq = ids.unique([(s, False), (d, False), (nbr, True)])
            idx = compact_frontier(nbr)  # non-padded slots: one compaction + gathers
            seeds = torch.cat([s, d])
            ei = torch.stack([ids.local(seeds.index_select(0, idx // k)),
                              ids.local(nbr.index_select(0, idx))]).long()
            opt.zero_grad(set_to_none=True)
            zz, lu = mem(uniq)
            z = enc(zz, lu, ei, hop.nbr_edge_time[r0:r1].reshape(-1).index_select(0, idx),
                    hop.nbr_edge_x[r0:r1].reshape(-1, D).index_select(0, idx))
            i_s, i_d = ids.local(s).long(), ids.local(d).long()
            pos = decoder(torch.cat([z[i_s], z[i_d]], 1))
            neg = decoder(torch.cat([z[i_s], z[i_d.roll(1)]], 1))
            mem.update_state(s, d, t[b_lo:b_hi], x[b_lo:b_hi])  # before backward, as the example
            loss = bce(pos, torch.ones_like(pos)) + bce(neg, torch.zeros_like(neg))
            loss.backward()
            average_gradients(params)
            opt.step()
            touched[torch.cat([s, d]).long()] = True
        return loss

    @torch.no_grad()  # the inference step (mem is in train() mode for its update order only)
    def run_shard():
        """The per-batch model step of examples/linkproppred/tgn.py (without the decoder): sampled
        neighbourhoods -> unique nodes -> memory.forward -> GraphAttentionEmbedding over the
        (seed -> neighbour) edge list -> memory.update_state."""
        hop = csr.sample_window(lo, hi, [k])[0]  # neighbourhoods of every batch of the shard
        z = None
        for b_lo in range(lo, hi, bs):
            b_hi = min(b_lo + bs, hi)
            s, d = src[b_lo:b_hi], dst[b_lo:b_hi]
            r0, r1 = 2 * (b_lo - lo), 2 * (b_hi - lo)
            nbr = hop.nbr_nids[r0:r1].reshape(-1)
            ids = _BatchIdSet(N, dev)
            uniq = ids.unique([(s, False), (d, False), (nbr, True)])
            if a.embedding:
                idx = compact_frontier(nbr)  # non-padded slots: one compaction + gathers
                seeds = torch.cat([s, d])
                ei = torch.stack([ids.local(seeds.index_select(0, idx // k)),
                                  ids.local(nbr.index_select(0, idx))]).long()
                zz, lu = mem(uniq)
                z = enc(zz, lu, ei, hop.nbr_edge_time[r0:r1].reshape(-1).index_select(0, idx),
                        hop.nbr_edge_x[r0:r1].reshape(-1, D).index_select(0, idx))
            else:
                mem(torch.cat([s, d]).long())
            mem.update_state(s, d, t[b_lo:b_hi], x[b_lo:b_hi])
            touched[torch.cat([s, d]).long()] = True  # rows this shard's update_state wrote
        return z

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if a.train:
        run_shard = train_shard  # noqa: F811  same timing harness, training step per batch
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
            'row': 'config 4: TGN ' + ('TRAINING step (memory + embedding + decoder, backward, '
                                       'gradient all-reduce, Adam)' if a.train else
                                       'memory + attention embedding (model step without the decoder)'
                                       if a.embedding else 'memory') +
                   ', time-sharded, memory join at the shard boundary',
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
