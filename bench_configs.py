#!/usr/bin/env python
"""End-to-end pipelines for BASELINE.json configs[2], [3] and [4] (the 4-GPU time-sharded form of
configs[3], TGN memory with the shard join, is bench_tgn_shard.py; configs[1]/headline is bench.py).

    python bench_configs.py --config 3                      # TGAT, tgbl-wiki-shaped, 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29544 bench_configs.py --config 5      # DyGFormer, time-sharded

config 3  per loader batch of 200 edges: random negatives -> 2-hop recent-neighbour sampling
          k=[20,20] over [src|dst|neg] (windowed hook through DGDataLoader/HookManager, the
          drop-in API) -> TGAT forward (2 layers, 2 heads, time 100, embed 172) -> 600 embeddings.
config 4  the loop of examples/linkproppred/tgn.py on one GPU, through the drop-in API: random
          negatives -> recent neighbours k=10 over [src|dst|neg] -> device de-duplication ->
          TGNMemory.forward -> GraphAttentionEmbedding -> link decoder -> update_state.
config 5  per rank a time-range shard of a synthetic stream: one pre-sampled window (k=31 so the
          DyGFormer sequence is 32), then per batch of 200 edges a DyGFormer forward (patch 1,
          4x50 channels, 2 layers, 2 heads) -> 2x200 embeddings.  No collective on the data path.

Forward/evaluation pipelines; `--train` adds the backward pass (tgm_attn_backward /
tgm_gae_backward + tgm_tgn_backward / tgm_dyg_backward through autograd) and an Adam step; config 5 under torchrun averages the
gradients across the time shards with one NCCL all-reduce per step (data-parallel training).  One JSON line on rank 0; CUDA-event
times, max over ranks; the CPU oracle is timed beside it on a few batches (config 3 only: the
numpy DyGFormer oracle is timed in bench_rows.py).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tgm_b200 import (DGData, DGDataLoader, DGraph, HookManager,  # noqa: E402
                      RandomNegativeEdgeSamplerHook, RecencyCSR, RecencyNeighborHook)
from tgm_b200.core.storage import DeviceCOOStorage  # noqa: E402
from tgm_b200.nn import TGAT, DyGFormer  # noqa: E402
from tgm_b200.parallel import (average_gradients, max_over_ranks, shard_batches,  # noqa: E402
                               sum_over_ranks)


def config3(a, dev):
    from bench_rows import wiki_stream
    src, dst, t, x, N = wiki_stream()
    ei = torch.from_numpy(np.stack([src, dst], 1))
    dg = DGraph(DGData.from_raw(torch.from_numpy(t), ei, torch.from_numpy(x)), device=dev)
    bs, nn = 200, [20, 20]
    torch.manual_seed(1337)
    model = TGAT(node_dim=1, edge_dim=x.shape[1], time_dim=100, embed_dim=172, num_layers=2,
                 n_heads=2, dropout=0.0).to(dev)
    model = model.train() if a.train else model.eval()
    # link decoder of the example (examples/linkproppred/tgat.py: LinkPredictor): plain torch MLP,
    # not part of the hot path; only used by --train to give the backward pass a loss
    decoder = torch.nn.Sequential(torch.nn.Linear(2 * 172, 172), torch.nn.ReLU(),
                                  torch.nn.Linear(172, 1)).to(dev)
    if a.cuda_graph and not a.train:
        raise SystemExit('--cuda-graph captures the training step: add --train')
    opt = torch.optim.Adam(list(model.parameters()) + list(decoder.parameters()), lr=1e-4,
                           capturable=bool(a.cuda_graph))
    node_x = torch.randn(N, 1, device=dev)
    hm = HookManager(keys=['train'])
    hm.register('train', RandomNegativeEdgeSamplerHook(low=8227, high=N))
    hm.register('train', RecencyNeighborHook(
        num_nodes=N, num_nbrs=nn, seed_nodes_keys=['edge_src', 'edge_dst', 'neg'],
        seed_times_keys=['edge_time', 'edge_time', 'neg_time'], window_batches=a.window_batches,
        lazy_edge_x=bool(a.lazy_edge_x) and not a.train))
    nb = a.batches

    losses = []
    graph = {}  # --cuda-graph: static input buffers + the captured training step

    def model_inputs(batch):
        return [*batch.seed_nids, *batch.seed_times, *batch.nbr_nids, *batch.nbr_edge_x,
                *batch.nbr_edge_time]

    def static_step():
        st, L = graph['static'], len(nn)
        z = model(node_x, st[0:L], st[L:2 * L], st[2 * L:3 * L], st[3 * L:4 * L], st[4 * L:5 * L])
        n = bs
        zs, zd, zn = z[:n], z[n:2 * n], z[2 * n:]
        pos = decoder(torch.cat([zs, zd], 1))
        neg = decoder(torch.cat([zs, zn], 1))
        loss = torch.nn.functional.binary_cross_entropy_with_logits(pos, torch.ones_like(pos)) + \
            torch.nn.functional.binary_cross_entropy_with_logits(neg, torch.zeros_like(neg))
        loss.backward()
        opt.step()
        return z, loss

    def capture(batch):
        """Whole model step (TGAT forward + tgm_attn_backward + Adam) as one CUDA graph over static
        copies of the hook outputs; the loader and the hooks stay eager Python."""
        graph['static'] = [torch.empty_like(v) for v in model_inputs(batch)]
        for dst_, src_ in zip(graph['static'], model_inputs(batch)):
            dst_.copy_(src_)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                opt.zero_grad(set_to_none=True)
                static_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        opt.zero_grad(set_to_none=True)
        graph['g'] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph['g']):
            graph['out'] = static_step()

    def epoch():
        hm.reset_state()
        done = 0
        losses.clear()
        with hm.activate('train'):
            for batch in DGDataLoader(dg, batch_size=bs, hook_manager=hm):
                if a.cuda_graph and batch.edge_src.numel() == bs and batch.neg.numel() == bs:
                    if 'g' not in graph:
                        capture(batch)
                    for dst_, src_ in zip(graph['static'], model_inputs(batch)):
                        dst_.copy_(src_)
                    graph['g'].replay()
                    z = graph['out'][0]
                    losses.append(graph['out'][1].detach().clone())
                    done += 1
                    if done == nb:
                        break
                    continue
                if a.train:
                    opt.zero_grad(set_to_none=True)
                z = model(node_x, batch.seed_nids, batch.seed_times, batch.nbr_nids,
                          batch.nbr_edge_x, batch.nbr_edge_time)
                if a.train:  # the example's loss: BCE on (src,dst) positives vs (src,neg) negatives
                    n = batch.edge_src.numel()
                    zs, zd, zn = z[:n], z[n:2 * n], z[2 * n:]
                    pos = decoder(torch.cat([zs, zd], 1))
                    neg = decoder(torch.cat([zs, zn], 1))
                    loss = torch.nn.functional.binary_cross_entropy_with_logits(
                        pos, torch.ones_like(pos)) + \
                        torch.nn.functional.binary_cross_entropy_with_logits(
                            neg, torch.zeros_like(neg))
                    loss.backward()
                    opt.step()
                    losses.append(loss.detach())
                done += 1
                if done == nb:
                    break
        return z

    epoch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    z = epoch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # CPU: C ring sampler + numpy TGAT oracle on a few batches (same shapes, same weights)
    from oracle import nn_oracle
    from oracle.c_oracle import CRing
    p = {k_: v.detach().cpu().numpy() for k_, v in model.state_dict().items()}
    ring = CRing(N, nn, x.shape[1])
    npx = node_x.cpu().numpy()
    rng = np.random.default_rng(0)
    nbc = 3
    c0 = time.perf_counter()
    for b in range(nbc):
        lo, hi = b * bs, (b + 1) * bs
        neg = rng.integers(8227, N, hi - lo).astype(np.int32)
        seeds = np.concatenate([src[lo:hi], dst[lo:hi], neg])
        tq = np.concatenate([t[lo:hi]] * 3)
        hops = ring.hook_call(seeds, tq, src[lo:hi], dst[lo:hi], t[lo:hi], x[lo:hi])
        nn_oracle.tgat_forward(p, 2, 2, npx, [h[0] for h in hops], [h[1] for h in hops],
                               [h[2] for h in hops], [h[4] for h in hops], [h[3] for h in hops])
    cpu_ms = (time.perf_counter() - c0) / nbc * 1e3
    return {'row': 'config 3: TGAT on tgbl-wiki-shaped stream, 2-hop k=[20,20], bs=200 (+200 negatives)',
            'batches': nb, 'ms_per_batch': ms / nb, 'events_per_s': nb * bs / (ms * 1e-3),
            'embeddings_per_batch': int(z.shape[0]),
            'sampled_edges_per_s': nb * (600 * 20 + 12000 * 20) / (ms * 1e-3),
            'cpu_baseline': {'value': cpu_ms, 'unit': 'ms/batch', 'kind': 'port',
                             'cores': os.cpu_count(),
                             'sample': f'{nbc} batches: C ring sampler + numpy TGAT oracle'},
            'mode': ('train (forward + tgm_attn_backward + Adam step, dropout 0)' +
                     (', model step replayed as one CUDA graph' if graph else '')) if a.train else 'forward',
            'loss_first_last': [float(losses[0]), float(losses[-1])] if losses else None,
            'note': 'DGDataLoader + HookManager + windowed RecencyNeighborHook + TGAT.forward' +
                    (' + BCE loss on a torch MLP decoder + backward + Adam' if a.train else '') +
                    '; the CPU figure is forward only'}


def config4(a, dev):
    """examples/linkproppred/tgn.py:60-124 (train) / :127-190 (eval, without the ranking metric)."""
    from bench_rows import wiki_stream
    from tgm_b200 import DeduplicationHook
    from tgm_b200.hooks.dedup import compact_frontier
    from tgm_b200.nn import GraphAttentionEmbedding, TGNMemory
    src, dst, t, x, N = wiki_stream()
    t = np.arange(len(t), dtype=np.int64) * 17  # unique times: the TGN-memory parity domain
    ei = torch.from_numpy(np.stack([src, dst], 1))
    dg = DGraph(DGData.from_raw(torch.from_numpy(t), ei, torch.from_numpy(x)), device=dev)
    bs, k, D, M, TD, Z = 200, 10, x.shape[1], 100, 100, 100
    torch.manual_seed(1337)
    mem = TGNMemory(N, D, M, TD).to(dev)
    enc = GraphAttentionEmbedding(M, Z, D, mem.time_enc)
    enc.conv.dropout = 0.0  # the B200 path trains with dropout 0 (cannot follow the reference RNG)
    enc = enc.to(dev)
    decoder = torch.nn.Sequential(torch.nn.Linear(2 * Z, Z), torch.nn.ReLU(),
                                  torch.nn.Linear(Z, 1)).to(dev)  # the example's LinkPredictor
    params = {id(q): q for m_ in (mem, enc, decoder) for q in m_.parameters()}  # time_enc is shared
    opt = torch.optim.Adam(params.values(), lr=1e-4)
    for m_ in (mem, enc, decoder):
        m_.train(a.train)
    hm = HookManager(keys=['train'])
    hm.register('train', RandomNegativeEdgeSamplerHook(low=8227, high=N))
    hm.register('train', RecencyNeighborHook(
        num_nodes=N, num_nbrs=[k], seed_nodes_keys=['edge_src', 'edge_dst', 'neg'],
        seed_times_keys=['edge_time', 'edge_time', 'neg_time']))
    hm.register('train', DeduplicationHook(seed_nodes_keys=['neg', 'nbr_nids']))
    nb = a.batches
    losses = []
    bce = torch.nn.functional.binary_cross_entropy_with_logits

    def epoch():
        hm.reset_state()
        mem.reset_state()
        losses.clear()
        done = 0
        with hm.activate('train'):
            for batch in DGDataLoader(dg, batch_size=bs, hook_manager=hm):
                if a.train:
                    opt.zero_grad(set_to_none=True)
                # the sampled frontier without its padded slots: ONE compaction (tgm_frontier_compact,
                # one host read for the count) and four gathers, instead of the example's four
                # boolean-mask selections (each a select + count read-back)
                nbr = batch.nbr_nids[0].flatten()
                idx = compact_frontier(nbr)
                seeds = torch.cat([batch.edge_src, batch.edge_dst, batch.neg])
                eidx = torch.stack([batch.global_to_local(seeds.index_select(0, idx // k)),
                                    batch.global_to_local(nbr.index_select(0, idx))]).long()
                z, lu = mem(batch.unique_nids)
                z = enc(z, lu, eidx, batch.nbr_edge_time[0].flatten().index_select(0, idx),
                        batch.nbr_edge_x[0].flatten(0, -2).index_select(0, idx))
                i_s, i_d, i_n = (batch.global_to_local(v).long()
                                 for v in (batch.edge_src, batch.edge_dst, batch.neg))
                pos = decoder(torch.cat([z[i_s], z[i_d]], 1))
                neg = decoder(torch.cat([z[i_s], z[i_n]], 1))
                # update memory with the ground-truth state BEFORE backward, as the example does
                mem.update_state(batch.edge_src, batch.edge_dst, batch.edge_time, batch.edge_x)
                if a.train:
                    loss = bce(pos, torch.ones_like(pos)) + bce(neg, torch.zeros_like(neg))
                    loss.backward()
                    opt.step()
                    mem.detach()
                    losses.append(loss.detach())
                done += 1
                if done == nb:
                    break
        return z

    epoch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    z = epoch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {'row': 'config 4 on one GPU: TGN (memory 100 + attention embedding 100) on a tgbl-wiki-'
                   'shaped stream, k=10, bs=200 (+200 negatives), through the drop-in loader/hook API',
            'batches': nb, 'ms_per_batch': ms / nb, 'events_per_s': nb * bs / (ms * 1e-3),
            'unique_nodes_last_batch': int(z.shape[0]),
            'mode': 'train (forward + tgm_gae_backward + tgm_tgn_backward + Adam step, dropout 0)'
                    if a.train else 'forward',
            'loss_first_last': [float(losses[0]), float(losses[-1])] if losses else None,
            'note': 'negatives -> ring sampler -> device dedup -> TGNMemory.forward -> '
                    'GraphAttentionEmbedding -> torch MLP decoder -> update_state' +
                    (' -> BCE -> backward -> Adam' if a.train else '') +
                    '; launch/Python-bound at bs=200; the 4-GPU time-sharded form with the memory '
                    'join is bench_tgn_shard.py'}


def config5(a, dev, rank, world):
    E, N, D, bs, k = a.edges, a.nodes, 16, 200, 31
    g = torch.Generator(device=dev).manual_seed(0)
    src = torch.randint(0, N, (E,), generator=g, device=dev, dtype=torch.int32)
    dst = torch.randint(0, N, (E,), generator=g, device=dev, dtype=torch.int32)
    t = torch.sort(torch.randint(0, 2000, (E,), generator=g, device=dev))[0]
    x = torch.randn(E, D, generator=g, device=dev)
    store = DeviceCOOStorage.from_device_tensors(src, dst, t, x, N)
    csr = RecencyCSR(store, bs, colocate_x=True)
    torch.manual_seed(0)
    model = DyGFormer(node_feat_dim=128, edge_x_dim=D, time_feat_dim=100, channel_embedding_dim=50,
                      output_dim=172, patch_size=1, num_layers=2, num_heads=2,
                      max_input_sequence_length=k + 1, dropout=0.0).to(dev)
    model = model.train() if a.train else model.eval()
    decoder = torch.nn.Sequential(torch.nn.Linear(2 * 172, 172), torch.nn.ReLU(),
                                  torch.nn.Linear(172, 1)).to(dev)  # the example's link predictor
    train_params = list(model.parameters()) + list(decoder.parameters())
    opt = torch.optim.Adam(train_params, lr=1e-4)
    node_x = torch.randn(N, 128, generator=g, device=dev)
    shard = shard_batches(E, bs, rank, world)
    lo = shard.edge_lo + (shard.num_edges // 2 // bs) * bs  # mid-shard: populated histories
    hi = min(lo + a.batches * bs, shard.edge_hi)

    graphed = None
    if a.cuda_graph:
        # Whole-step CUDA graph (forward + tgm_dyg_backward + Adam) over static input buffers: the
        # step is ~250 launches behind ctypes/torch dispatch, i.e. launch-bound.  Every library
        # call is stream-ordered and allocation-free once its scratch has grown (the warm-up
        # steps), and tgm_dyg_set_params -- device-to-device copies from the parameter tensors --
        # is captured with the forward, so the handle's weight copies follow the in-graph Adam.
        if not a.train or world > 1 or (hi - lo) % bs:
            raise SystemExit('--cuda-graph: single-GPU --train over whole batches only')
        opt = torch.optim.Adam(train_params, lr=1e-4, capturable=True)
        st_ei = torch.empty((2, bs), dtype=torch.int32, device=dev)
        st_t = torch.empty((bs,), dtype=torch.int64, device=dev)
        st_nb = torch.empty((2 * bs, k), dtype=torch.int32, device=dev)
        st_nt = torch.empty((2 * bs, k), dtype=torch.int64, device=dev)
        st_nx = torch.empty((2 * bs, k, D), dtype=torch.float32, device=dev)

        def load(hop, b_lo):
            r0 = 2 * (b_lo - lo)
            st_ei[0].copy_(src[b_lo:b_lo + bs])
            st_ei[1].copy_(dst[b_lo:b_lo + bs])
            st_t.copy_(t[b_lo:b_lo + bs])
            st_nb.copy_(hop.nbr_nids[r0:r0 + 2 * bs])
            st_nt.copy_(hop.nbr_edge_time[r0:r0 + 2 * bs])
            st_nx.copy_(hop.nbr_edge_x[r0:r0 + 2 * bs])

        def static_step():
            zs, zd = model(node_x, st_ei, st_t, st_nb, st_nt, st_nx)
            logit = decoder(torch.cat([zs, zd], 1))
            loss = torch.nn.functional.binary_cross_entropy_with_logits(logit, torch.ones_like(logit))
            loss.backward()
            opt.step()
            return zs

        hop0 = csr.sample_window(lo, lo + bs, [k])[0]
        load(hop0, lo)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):  # grows every scratch buffer, builds cuBLAS/optimizer state
                opt.zero_grad(set_to_none=True)
                static_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        opt.zero_grad(set_to_none=True)
        graphed = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graphed):
            st_zs = static_step()

    def run():
        hop = csr.sample_window(lo, hi, [k])[0]
        if graphed is not None:
            for b_lo in range(lo, hi, bs):
                load(hop, b_lo)
                graphed.replay()
            return st_zs
        for b_lo in range(lo, hi, bs):
            b_hi = min(b_lo + bs, hi)
            r0, r1 = 2 * (b_lo - lo), 2 * (b_hi - lo)
            ei = torch.stack([src[b_lo:b_hi], dst[b_lo:b_hi]])
            if a.train:
                opt.zero_grad(set_to_none=True)
            zs, zd = model(node_x, ei, t[b_lo:b_hi], hop.nbr_nids[r0:r1], hop.nbr_edge_time[r0:r1],
                           hop.nbr_edge_x[r0:r1])
            if a.train:  # positives only: the backward pass and the optimizer step are what is timed
                logit = decoder(torch.cat([zs, zd], 1))
                loss = torch.nn.functional.binary_cross_entropy_with_logits(logit, torch.ones_like(logit))
                loss.backward()
                if world > 1:  # data-parallel over the time shards: average the gradients (NCCL)
                    average_gradients(train_params)
                opt.step()
        return zs

    run()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), dev)
    events = sum_over_ranks(float(hi - lo), dev)
    nb = (hi - lo) // bs
    return {'row': 'config 5: DyGFormer on a time-sharded synthetic stream, sequence 32 (k=31), patch 1',
            'n_gpus': world, 'edges': E, 'batches_per_rank': nb, 'ms_per_batch': ms / nb,
            'events_per_s': events / (ms * 1e-3), 'sequences_per_s': 2 * events / (ms * 1e-3),
            'mode': ('train (forward + tgm_dyg_backward + Adam step, dropout 0)' +
                     (', whole step replayed as one CUDA graph' if graphed is not None else ''))
                    if a.train else 'forward',
            'note': 'one pre-sampled window per rank + DyGFormer.forward per 200-edge batch' +
                    (' + BCE loss on a torch MLP decoder + backward + gradient all-reduce (NCCL, when '
                     'n_gpus > 1) + Adam' if a.train else '') +
                    '; store/adjacency replicated, no collective on the sampling path'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--train', action='store_true', help='forward + backward + Adam step per batch')
    ap.add_argument('--config', type=int, required=True, choices=[3, 4, 5])
    ap.add_argument('--cuda-graph', action='store_true',
                    help='configs 3 and 5 with --train on one GPU: capture the training step (forward + '
                         'backward + Adam) in a CUDA graph and replay it per batch')
    ap.add_argument('--batches', type=int, default=200)
    ap.add_argument('--window-batches', type=int, default=25)
    ap.add_argument('--lazy-edge-x', type=int, default=1,
                    help='config 3 inference: the sampler hands out edge ids and the attention reads '
                         'the feature rows in place (1) instead of materialising nbr_edge_x (0)')
    ap.add_argument('--edges', type=int, default=100_000_000)
    ap.add_argument('--nodes', type=int, default=1_000_000)
    a = ap.parse_args()
    torch.set_grad_enabled(bool(a.train))  # forward pipelines unless --train
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=dev)
    out = (config3(a, dev) if a.config == 3 else config4(a, dev) if a.config == 4
           else config5(a, dev, rank, world))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
