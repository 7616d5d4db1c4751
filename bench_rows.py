#!/usr/bin/env python
"""Per-row measurements for SURVEY.md section 8a (everything except the headline sampler, which
bench.py owns): each row is timed on the B200 with CUDA events after warm-up, reported against the
roofline that bounds it (measured HBM copy peak from MEASURED_PEAKS.json), with the CPU oracle
timed beside it on the host on a bounded sample.  One JSON line per row.

    python bench_rows.py [--rows store,ring,...] > profiles/rNN_rows.jsonl

Rows with a model in them also carry `eager_cuda_baseline`: the reference's own tensor-op sequence
(oracle/torch_eager.py, pinned on the reference fixtures) run on the SAME B200 -- what the
reference's device='cuda' mode launches -- so the ratio is GPU over GPU, not GPU over numpy.

oracle/ is executed here only as a timed baseline, never on the product path.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tgm_b200 import (DGData, DGDataLoader, DGraph, HookManager, NeighborSamplerHook,  # noqa: E402
                      RecencyCSR, RecencyNeighborHook, _cabi)
from tgm_b200.core.storage import DGSliceTracker  # noqa: E402
from tgm_b200.hooks.dedup import compact_frontier  # noqa: E402
from tgm_b200.nn import TGAT, DyGFormer, TemporalAttention, TGNMemory, Time2Vec, masked_mean  # noqa: E402

DEV = torch.device('cuda', 0)
try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    PEAK_SRC = 'measured'
except Exception:  # noqa: BLE001
    PEAK, PEAK_SRC = 6650.0, 'fallback'


def cuda_ms(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def cpu_s(fn, iters=1):
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters


def hbm(bytes_, ms):
    ach = bytes_ / (ms * 1e-3) / 1e9
    return {'bound': 'hbm', 'achieved': ach, 'peak': PEAK, 'peak_source': PEAK_SRC, 'unit': 'GB/s',
            'frac': ach / PEAK, 'algorithmic_bytes': bytes_}


def emit(row, **kw):
    print(json.dumps({'row': row, **kw}), flush=True)


def wiki_stream(seed=0, E=157_474, N=9227, D=172, T=2_678_374):
    """tgbl-wiki-shaped synthetic stand-in (bipartite 8227 users x 1000 items)."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, 8227, E).astype(np.int32)
    dst = rng.integers(8227, N, E).astype(np.int32)
    t = np.sort(rng.integers(0, T, E)).astype(np.int64)
    x = rng.standard_normal((E, D)).astype(np.float32)
    return src, dst, t, x, N


def graph(src, dst, t, x):
    ei = torch.from_numpy(np.stack([src, dst], 1))
    return DGraph(DGData.from_raw(torch.from_numpy(t), ei, None if x is None else torch.from_numpy(x)),
                  device=DEV)


# ---- S1-S4: slice + materialise one loader batch -------------------------------------------------
def row_store():
    E, N, D, bs = 10_000_000, 1_000_000, 16, 200
    g = torch.Generator(device=DEV).manual_seed(0)
    src = torch.randint(0, N, (E,), generator=g, device=DEV, dtype=torch.int32)
    dst = torch.randint(0, N, (E,), generator=g, device=DEV, dtype=torch.int32)
    t = torch.sort(torch.randint(0, 2000, (E,), generator=g, device=DEV))[0]
    x = torch.randn(E, D, generator=g, device=DEV)
    data = DGData.from_raw(t.cpu(), torch.stack([src, dst], 1).cpu(), x.cpu())
    dg = DGraph(data, device=DEV)
    starts = list(range(0, E - bs, E // 2000))

    def materialise():
        for s in starts[:200]:
            dg.slice_events(s, s + bs).materialize()
    t_gpu = cpu_s(materialise) / 200  # host-side work only: the slab is a zero-copy device view
    # reference algorithm (array_backend.py:57-68, 259-268): two O(E) boolean masks per batch
    em = np.arange(E, dtype=np.int32)
    ei, tt, xx = data.edge_index.numpy(), data.time.numpy(), data.edge_x.numpy()

    def ref_batch(s=E // 2):
        m = (em >= s) & (em < s + bs)
        a, b = ei[m], tt[m]
        m2 = (em >= s) & (em < s + bs)
        return a, b, xx[m2]
    t_cpu = cpu_s(ref_batch, 5)
    emit('S1-S4 slice+materialize one batch (E=1e7, bs=200, D=16)', value=t_gpu * 1e6, unit='us/batch',
         higher_is_better=False,
         note='two binary searches on a host mirror + pointer offsets; no kernel, no H2D',
         cpu_baseline={'value': t_cpu * 1e6, 'unit': 'us/batch', 'kind': 'port', 'cores': 1,
                       'sample': 'numpy restatement of the two O(E) masks of get_edges/get_edge_x'})


# ---- R1-R4 stateful ring through the drop-in loader + hook API (config 1 stand-in) ---------------
def row_ring():
    from oracle.c_oracle import CRing
    src, dst, t, x, N = wiki_stream()
    dg = graph(src, dst, t, x)
    bs, k = 200, 10
    hm = HookManager(keys=['g'])
    hm.register('g', RecencyNeighborHook(num_nodes=N, num_nbrs=[k],
                                         seed_nodes_keys=['edge_src', 'edge_dst'],
                                         seed_times_keys=['edge_time', 'edge_time']))
    nb = 0
    with hm.activate('g'):
        for batch in DGDataLoader(dg, batch_size=bs, hook_manager=hm):  # warm-up epoch
            nb += 1
        # an epoch is ~8 ms of wall clock: one allocator or scheduler hiccup shows as tens of
        # us per batch (fresh-box runs ranged 8.4 .. 59), so five epochs are timed and the
        # fastest reported, all of them listed
        epochs = []
        for _ in range(5):
            hm.reset_state()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for batch in DGDataLoader(dg, batch_size=bs, hook_manager=hm):
                pass
            torch.cuda.synchronize()
            epochs.append(time.perf_counter() - t0)
        dt = min(epochs)
    slots = 2 * len(src) * k
    # the same epoch on the stateful ring kernels, batch by batch (window_batches=0)
    hm0 = HookManager(keys=['g'])
    hm0.register('g', RecencyNeighborHook(num_nodes=N, num_nbrs=[k],
                                          seed_nodes_keys=['edge_src', 'edge_dst'],
                                          seed_times_keys=['edge_time', 'edge_time'],
                                          window_batches=0))
    with hm0.activate('g'):
        for batch in DGDataLoader(dg, batch_size=bs, hook_manager=hm0):
            pass
        ring_epochs = []
        for _ in range(3):
            hm0.reset_state()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for batch in DGDataLoader(dg, batch_size=bs, hook_manager=hm0):
                pass
            torch.cuda.synchronize()
            ring_epochs.append(time.perf_counter() - t0)
        dt_ring = min(ring_epochs)
    ring = CRing(N, [k], x.shape[1])
    c0 = time.perf_counter()
    ring.run_stream(src, dst, t, x, 0, len(src), bs, checksum=False)
    cdt = time.perf_counter() - c0
    # the reference's tensor-op sequence on this GPU (hook only, batches pre-sliced)
    from oracle.torch_eager import TorchRing
    tr = TorchRing(N, [k], x.shape[1], device=DEV)
    dsrc, ddst, dt_, dx = (torch.from_numpy(v).to(DEV) for v in (src, dst, t, x))

    def eager_epoch(nb_):
        for b in range(nb_):
            lo, hi = b * bs, min((b + 1) * bs, len(src))
            tr.hook_call(torch.cat([dsrc[lo:hi], ddst[lo:hi]]), torch.cat([dt_[lo:hi], dt_[lo:hi]]),
                         dsrc[lo:hi], ddst[lo:hi], dt_[lo:hi], dx[lo:hi])
        torch.cuda.synchronize()
    eager_epoch(50)
    tr.reset_state()
    e0 = time.perf_counter()
    eager_epoch(nb)
    edt = time.perf_counter() - e0
    emit('R1-R4 DGDataLoader + RecencyNeighborHook epoch (wiki-shaped, bs=200, k=10, D=172)',
         value=slots / dt, unit='sampled-edges/s', batches_per_s=nb / dt, us_per_batch=dt / nb * 1e6,
         ring_kernels_us_per_batch=dt_ring / nb * 1e6,
         us_per_batch_all_epochs=[round(e / nb * 1e6, 2) for e in epochs],
         ring_kernels_us_per_batch_all_epochs=[round(e / nb * 1e6, 2) for e in ring_epochs],
         note='fastest of the timed epochs (all listed); default-constructed hook: windows of 1024 batches pre-sampled by one launch per hop, '
              'each call hands out views (Python-bound); ring_kernels_us_per_batch = the same epoch '
              'with window_batches=0 (one tgm_recency_step call = 3 launches per batch)',
         cpu_baseline={'value': slots / cdt, 'unit': 'sampled-edges/s', 'kind': 'port', 'cores': 1,
                       'sample': 'C port of the ring sampler over the same epoch'},
         eager_cuda_baseline={'value': slots / edt, 'unit': 'sampled-edges/s',
                              'us_per_batch': edt / nb * 1e6,
                              'sample': 'oracle/torch_eager.py::TorchRing over the same epoch on '
                                        'cuda:0, hook only (batches pre-sliced on the device)'})
    # the ring kernels alone at a size where bandwidth shows: 1M random seeds against full rings
    N2, B, D = 1_000_000, 20, 16
    h = RecencyNeighborHook(N2, [B], ['edge_src'], ['edge_time'])
    g = torch.Generator(device=DEV).manual_seed(1)

    class _G:  # minimal dg stand-in for state creation
        device, edge_x_dim = DEV, D
    h._ensure_state(_G)
    stream = _cabi.current_stream(DEV)
    # full rings, written straight into the live state (tgm_recency_state): what 20+ pushes per
    # node leave behind
    import ctypes
    ptrs = [ctypes.c_void_p() for _ in range(4)]
    _cabi.check(_cabi.lib.tgm_recency_state(h._handle, *[ctypes.byref(q) for q in ptrs]))
    _cabi.device_view(ptrs[0].value, (N2, B), torch.int32, DEV).copy_(
        torch.randint(0, N2, (N2, B), generator=g, device=DEV, dtype=torch.int32))
    _cabi.device_view(ptrs[1].value, (N2, B), torch.int64, DEV).copy_(
        torch.sort(torch.randint(0, 10 ** 6, (N2, B), generator=g, device=DEV), 1)[0])
    _cabi.device_view(ptrs[2].value, (N2, B, D), torch.float32, DEV).normal_(generator=g)
    _cabi.device_view(ptrs[3].value, (N2,), torch.int32, DEV).fill_(B)
    S = 1_000_000
    seeds = torch.randint(0, N2, (S,), generator=g, device=DEV, dtype=torch.int32)
    tq = torch.full((S,), 10 ** 9, device=DEV, dtype=torch.int64)
    ms = cuda_ms(lambda: h._query(seeds, tq, B, stream))
    bytes_ = S * (B * 12 + 2 * B * (12 + 4 * D) + 16)
    emit('R1 ring_query_kernel (N=1e6 rings, B=k=20, D=16, 1e6 random seeds)', value=S * B / (ms * 1e-3),
         unit='sampled-edges/s', ms=ms, roofline=hbm(bytes_, ms))


# ---- R1 stateless, general seeds: 2-hop window (config 3 sampling) --------------------------------
def row_twohop():
    src, dst, t, x, N = wiki_stream()
    dg = graph(src, dst, t, x)
    bs, nn = 200, [20, 20]
    csr = RecencyCSR(dg._storage, bs, colocate_x=True)
    nbatch = 25
    lo, hi = 100_000, 100_000 + nbatch * bs
    neg = torch.randint(8227, N, (hi - lo,), device=DEV, dtype=torch.int32)
    ms = cuda_ms(lambda: csr.sample_window(lo, hi, nn, neg=neg), iters=5)
    S0 = 3 * (hi - lo)
    slots = S0 * 20 + S0 * 20 * 20
    D = x.shape[1]
    emit('R1 csr_sample_fast_kernel 2-hop window (wiki-shaped, 25 batches x 600 seeds, k=[20,20], D=172)',
         value=slots / (ms * 1e-3), unit='sampled-edges/s', ms=ms,
         roofline=hbm(slots * 2 * (12 + 4 * D), ms),
         note='includes the torch-side seed layout of the window; hop-1 features 4 GB per window')


# ---- R1 stateless hop 0: the forms without the feature block, device-resident ---------------------
def row_forms():
    """tgm_csr_sample_edges_ids / _mean on a stream with the headline workload's statistics at a
    tenth of its size (E=1e7, N=1e5: the same 100 entries per node, D=16, k=B=20), windows of 5000
    batches from the steady state, outputs preallocated; a different window every call (>> L2)."""
    E, N, D, k, bs, W = 10_000_000, 100_000, 16, 20, 200, 5000
    g = torch.Generator(device=DEV).manual_seed(0)
    src = torch.randint(0, N, (E,), generator=g, device=DEV, dtype=torch.int32)
    dst = torch.randint(0, N, (E,), generator=g, device=DEV, dtype=torch.int32)
    t = torch.sort(torch.randint(0, 2000, (E,), generator=g, device=DEV))[0]
    x = torch.randn(E, D, generator=g, device=DEV)
    from tgm_b200.core.storage import DeviceCOOStorage
    st = DeviceCOOStorage.from_device_tensors(src, dst, t, x, N)
    csr = RecencyCSR(st, bs, colocate_x=True)
    me = W * bs
    wins = [(lo, lo + me) for lo in range(3_000_000, E - me + 1, me)]
    S = 2 * me
    state = {'i': 0}

    def window():
        state['i'] += 1
        return wins[state['i'] % len(wins)]
    out_ids = (torch.empty((S, k), dtype=torch.int32, device=DEV), torch.empty((S, k), dtype=torch.int64, device=DEV),
               torch.empty((S, k), dtype=torch.int32, device=DEV))
    out_mean = (None, None, torch.empty((S, D), dtype=torch.float32, device=DEV))
    out_full = (out_ids[0], out_ids[1], torch.empty((S, k, D), dtype=torch.float32, device=DEV))
    ms_full = cuda_ms(lambda: csr.sample_edges(*window(), k, k, out=out_full), iters=6)
    ms_ids = cuda_ms(lambda: csr.sample_edges_ids(*window(), k, k, out=out_ids), iters=6)
    ms_mean = cuda_ms(lambda: csr.sample_edges_mean(*window(), k, k, with_ids=False, out=out_mean), iters=6)
    slots = S * k
    emit('R1 hop-0 forms, device-resident (E=1e7, N=1e5, D=16, k=B=20, 5000-batch windows)',
         value=slots / (ms_ids * 1e-3), unit='sampled-edges/s (id form)', ms=ms_ids,
         roofline=hbm(slots * (16 + 16) + S * 20, ms_ids),
         full_rows={'ms': ms_full, 'sampled_edges_per_s': slots / (ms_full * 1e-3),
                    'roofline': hbm(slots * 2 * (12 + 4 * D) + S * 20, ms_full)},
         fused_mean={'ms': ms_mean, 'sampled_edges_per_s': slots / (ms_mean * 1e-3),
                     'roofline': hbm(slots * (12 + 4 * D) + S * (4 * D + 20), ms_mean)},
         note='id form: 16 B entry read + (nid, t, eid) = 16 B written per slot; fused mean: entry '
              'ids/times + feature rows read (12 + 4D per slot), 4D written per SEED; full rows: the '
              'headline kernel on this smaller stream')


# ---- S5 uniform full-history sampler ---------------------------------------------------------------
def row_uniform():
    E, N, D, k = 10_000_000, 1_000_000, 16, 20
    g = torch.Generator(device=DEV).manual_seed(0)
    src = torch.randint(0, N, (E,), generator=g, device=DEV, dtype=torch.int32)
    dst = torch.randint(0, N, (E,), generator=g, device=DEV, dtype=torch.int32)
    t = torch.sort(torch.randint(0, 2000, (E,), generator=g, device=DEV))[0]
    x = torch.randn(E, D, generator=g, device=DEV)
    from tgm_b200.core.storage import DeviceCOOStorage
    st = DeviceCOOStorage.from_device_tensors(src, dst, t, x, N)
    seeds = torch.randint(0, N, (600,), generator=g, device=DEV, dtype=torch.int32)
    sl = DGSliceTracker(end_time=1500)
    st.get_nbrs(seeds, k, sl, False)  # builds the (edge, side) adjacency once
    ms = cuda_ms(lambda: st.get_nbrs(seeds, k, sl, False, ), iters=20)
    big = torch.randint(0, N, (1_000_000,), generator=g, device=DEV, dtype=torch.int32)
    ms_big = cuda_ms(lambda: st.get_nbrs(big, k, sl, False), iters=5)
    # reference algorithm on the host: one pass over every edge of the history per call
    n_hist = 200_000
    s_np, d_np = src[:n_hist].cpu().numpy().tolist(), dst[:n_hist].cpu().numpy().tolist()
    want = set(seeds.cpu().numpy().tolist())

    def ref_pass():
        nb = {v: [] for v in want}
        for e, (a, b) in enumerate(zip(s_np, d_np)):
            if a in nb:
                nb[a].append((e, b))
            if b in nb:
                nb[b].append((e, a))
    t_cpu = cpu_s(ref_pass) * (7_500_000 / n_hist)
    emit('S5 get_nbrs uniform (E=1e7 history, 600 seeds, k=20, D=16)', value=ms * 1e3, unit='us/call',
         higher_is_better=False, ms_1e6_seeds=ms_big,
         sampled_edges_per_s_1e6_seeds=1e6 * k / (ms_big * 1e-3),
         cpu_baseline={'value': t_cpu * 1e6, 'unit': 'us/call', 'kind': 'port', 'cores': 1,
                       'sample': 'pure-Python candidate pass of get_nbrs over 2e5 edges, scaled '
                                 'linearly to the 7.5e6-edge history the call covers'})


# ---- A1/A4 + frontier: streaming kernels ---------------------------------------------------------
def row_stream_kernels():
    S, k, D = 2_000_000, 20, 16
    g = torch.Generator(device=DEV).manual_seed(0)
    z = torch.randn(S, k, D, generator=g, device=DEV)
    nid = torch.randint(-1, 1000, (S, k), generator=g, device=DEV, dtype=torch.int32)
    ms = cuda_ms(lambda: masked_mean(z, nid))
    emit('A4 masked_mean_kernel (2e6 seeds, k=20, D=16)', value=S / (ms * 1e-3), unit='seeds/s', ms=ms,
         roofline=hbm(S * (k * D * 4 + k * 4 + D * 4), ms))
    flat = nid.reshape(-1)
    from tgm_b200 import _cabi
    st = torch.cuda.current_stream(DEV).cuda_stream
    idx = torch.empty(flat.numel(), dtype=torch.int64, device=DEV)
    cnt = torch.zeros(1, dtype=torch.int64, device=DEV)
    ms = cuda_ms(lambda: _cabi.check(_cabi.lib.tgm_frontier_compact(
        flat.data_ptr(), flat.numel(), idx.data_ptr(), cnt.data_ptr(), st)))
    kept = int(cnt.item())
    ms_api = cuda_ms(lambda: compact_frontier(flat))
    emit('frontier compaction (4e7 slots)', value=flat.numel() / (ms * 1e-3), unit='slots/s', ms=ms,
         roofline=hbm(flat.numel() * 4 + kept * 8, ms), kept=kept, ms_python_api_incl_count_sync=ms_api,
         note='one pass over the ids (segment masks in shared memory, one exchange of segment '
              'totals); algorithmic bytes = 4 per slot read + 8 per kept slot written')
    te = Time2Vec(100).to(DEV)
    dt = torch.randint(0, 2_600_000, (400_000,), generator=g, device=DEV)
    from tgm_b200 import _cabi
    w, b = te.w.weight.detach().reshape(-1).contiguous(), te.w.bias.detach().contiguous()
    out = torch.empty((dt.numel(), 100), device=DEV)
    st = torch.cuda.current_stream(DEV).cuda_stream
    ms = cuda_ms(lambda: _cabi.check(_cabi.lib.tgm_time2vec(dt.data_ptr(), dt.numel(), w.data_ptr(),
                                                            b.data_ptr(), 100, out.data_ptr(), st)))
    with torch.no_grad():
        ms_module = cuda_ms(lambda: te(dt))
    emit('A1 time2vec_kernel (4e5 deltas x 100 dims)', value=dt.numel() * 100 / (ms * 1e-3),
         unit='encodings/s', ms=ms, ms_through_the_module=ms_module,
         roofline=hbm(dt.numel() * (8 + 400), ms),
         note='tgm_time2vec into a preallocated output: warp per row, 15-instruction cosine (1.6e-7 '
              'abs); issue-bound (IPC 2.2 per SM, fixed-latency stalls of the polynomial chain). A '
              'variant with no idle lanes (8 rows = 25 whole tiles per warp) measured the same '
              '84 us and was dropped')


# ---- A2/A3 TGAT (config 3) --------------------------------------------------------------------------
def row_tgat():
    from oracle import nn_oracle
    rng = np.random.default_rng(0)
    torch.manual_seed(0)
    N, D, TD, EMB, k, S0 = 9227, 172, 100, 172, 20, 600
    model = TGAT(node_dim=1, edge_dim=D, time_dim=TD, embed_dim=EMB, num_layers=2, n_heads=2).to(DEV).eval()
    node_x = torch.randn(N, 1, device=DEV)
    sizes = [S0, S0 * k]
    hop = {}
    for h, S in enumerate(sizes):
        nid = rng.integers(0, N, (S, k)).astype(np.int32)
        pad = np.arange(k)[None, :] < rng.integers(0, k + 1, S)[:, None]
        nid[pad] = -1
        st = rng.integers(100_000, 2_600_000, S)
        nt = np.sort(np.clip(st[:, None] - rng.integers(1, 90_000, (S, k)), 0, None), 1)
        nt[pad] = 0
        ex = rng.standard_normal((S, k, D)).astype(np.float32)
        ex[pad] = 0
        seeds = rng.integers(0, N, S).astype(np.int32) if h == 0 else hop[0][2].reshape(-1)
        stt = st if h == 0 else hop[0][3].reshape(-1)
        hop[h] = (seeds, stt, nid, nt, ex)
    dv = lambda i: [torch.from_numpy(np.ascontiguousarray(hop[h][i])).to(DEV) for h in range(2)]
    args = (node_x, dv(0), dv(1), dv(2), dv(4), dv(3))
    torch.set_grad_enabled(False)  # inference rows: keep autograd out of the timings
    ms = cuda_ms(lambda: model(*args), iters=10)
    att = model.attn[0]
    a1 = (model.time_encoder, torch.randn(sizes[1], 1, device=DEV), torch.randn(sizes[1], k, 1, device=DEV),
          args[4][1], args[2][1], args[5][1], args[3][1])
    ms_att = cuda_ms(lambda: att.forward_fused(*a1), iters=10)
    p = {k_: v.detach().cpu().numpy() for k_, v in model.state_dict().items()}
    npx = node_x.cpu().numpy()
    t_cpu = cpu_s(lambda: nn_oracle.tgat_forward(p, 2, 2, npx, [hop[h][0] for h in range(2)],
                                                 [hop[h][1] for h in range(2)], [hop[h][2] for h in range(2)],
                                                 [hop[h][4] for h in range(2)], [hop[h][3] for h in range(2)]))
    from oracle import torch_eager as te
    pt = {k_: v.detach() for k_, v in model.state_dict().items()}
    largs = [[a_.long() if a_.dtype == torch.int32 else a_ for a_ in lst] for lst in args[1:]]
    eager_fwd = lambda: te.tgat_forward(pt, 2, 2, node_x, *largs)  # noqa: E731
    err = float((eager_fwd() - model(*args)).abs().max())
    ms_eager = cuda_ms(eager_fwd, iters=10)
    key1, out1 = 1 + D + TD, 102
    flops_ref = 2 * sizes[1] * k * key1 * 2 * out1
    emit('A2/A3 TGAT forward, one batch (600 seeds, k=[20,20], edge 172, time 100, embed 172)',
         value=ms, unit='ms/batch', higher_is_better=False,
         layer1_attention_ms_12000x20=ms_att,
         layer1_attention_roofline=hbm(sizes[1] * (k * (1 + D) * 4 + k * 12 + 2 * 2 * key1 * 4), ms_att),
         note=f'reassociated single-query attention: the reference W_KV GEMM ({flops_ref / 1e9:.1f} '
              'GFLOP per layer-1 call) is replaced by two S-row skinny GEMMs; the neighbour pass is '
              'SIMT fp32 (Time2Vec cosines + dot products over cp.async-staged rows), '
              'issue/latency-bound rather than HBM-bound',
         cpu_baseline={'value': t_cpu * 1e3, 'unit': 'ms/batch', 'kind': 'port', 'cores': os.cpu_count(),
                       'sample': 'numpy oracle of the same batch (BLAS threads = all cores)'},
         eager_cuda_baseline={'value': ms_eager, 'unit': 'ms/batch', 'max_abs_diff_vs_b200_path': err,
                              'sample': 'oracle/torch_eager.py::tgat_forward (the reference module\'s '
                                        'op sequence: cat -> W_KV GEMM over S*k rows -> softmax -> '
                                        'W_O -> LayerNorm) on cuda:0, same batch, same weights'})


# ---- A6 TGN memory (config 4) ---------------------------------------------------------------------
def row_tgn():
    from oracle.tgn_oracle import TGNMemoryOracle
    N, D, M, TD, bs = 1_000_000, 16, 100, 100, 200
    torch.manual_seed(0)
    mem = TGNMemory(N, D, M, TD).to(DEV)
    mem.train()
    mem.reset_state()
    g = torch.Generator(device=DEV).manual_seed(0)
    nb = 200
    E = nb * bs
    src = torch.randint(0, N, (E,), generator=g, device=DEV, dtype=torch.int32)
    dst = torch.randint(0, N, (E,), generator=g, device=DEV, dtype=torch.int32)
    t = torch.arange(E, device=DEV, dtype=torch.int64) * 7
    x = torch.randn(E, D, generator=g, device=DEV)
    nids = [torch.cat([src[i * bs:(i + 1) * bs], dst[i * bs:(i + 1) * bs]]).long() for i in range(nb)]

    @torch.no_grad()  # the inference row: no autograd rows are saved
    def epoch():
        for i in range(nb):
            lo, hi = i * bs, (i + 1) * bs
            mem(nids[i])
            mem.update_state(src[lo:hi], dst[lo:hi], t[lo:hi], x[lo:hi])
    ms = cuda_ms(epoch, warmup=1, iters=3) / nb
    No = 20_000  # the oracle's per-node dict store makes N=1e6 impractical on the host
    p = {k_: v.detach().cpu().numpy() for k_, v in mem.state_dict().items()}
    orc = TGNMemoryOracle(No, D, M, TD, p)
    s_np, d_np = (src.cpu().numpy() % No), (dst.cpu().numpy() % No)
    t_np, x_np = t.cpu().numpy(), x.cpu().numpy()

    def cpu_epoch():
        for i in range(20):
            lo, hi = i * bs, (i + 1) * bs
            orc.forward(np.unique(np.concatenate([s_np[lo:hi], d_np[lo:hi]])))
            orc.update_state(s_np[lo:hi], d_np[lo:hi], t_np[lo:hi], x_np[lo:hi])
    t_cpu = cpu_s(cpu_epoch) / 20
    from oracle.torch_eager import TorchTGNMemory
    pt = {k_: v.detach() for k_, v in mem.state_dict().items()}
    tmem = TorchTGNMemory(N, D, M, TD, pt, device=DEV)  # builds the 2 x N-entry dict store (tgn.py:183)
    sl_, dl_ = src.long(), dst.long()

    def eager_epoch(nb_=50):
        for i in range(nb_):
            lo, hi = i * bs, (i + 1) * bs
            tmem.forward(nids[i].unique())
            tmem.update_state(sl_[lo:hi], dl_[lo:hi], t[lo:hi], x[lo:hi])
        torch.cuda.synchronize()
    eager_epoch(10)
    e0 = time.perf_counter()
    eager_epoch(50)
    ms_eager = (time.perf_counter() - e0) / 50 * 1e3
    emit('A6 TGNMemory forward + update_state per batch (N=1e6, bs=200, D=16, M=100)', value=ms,
         unit='ms/batch', higher_is_better=False,
         note='11 launches per batch (gather/message, 2 SGEMM, GRU gates, scatter, store) x2; '
              'launch-latency bound at bs=200',
         cpu_baseline={'value': t_cpu * 1e3, 'unit': 'ms/batch', 'kind': 'port', 'cores': os.cpu_count(),
                       'sample': 'numpy oracle, 20 batches, N=2e4 nodes'},
         eager_cuda_baseline={'value': ms_eager, 'unit': 'ms/batch',
                              'sample': 'oracle/torch_eager.py::TorchTGNMemory on cuda:0, N=1e6, 50 '
                                        'batches (per-node Python dict message store as upstream)'})


# ---- N4 full TGN inference step through the drop-in API (examples/linkproppred/tgn.py loop) --------
def row_tgn_step():
    from oracle.tgn_oracle import TGNMemoryOracle, graph_attention_embedding
    from tgm_b200 import DeduplicationHook, RandomNegativeEdgeSamplerHook
    from tgm_b200.nn import GraphAttentionEmbedding
    src, dst, t, x, N = wiki_stream()
    nb, bs, k, D, M, TD = 300, 200, 10, 172, 100, 100
    E = nb * bs
    t = np.arange(len(t), dtype=np.int64) * 17  # unique times: the TGN-memory parity domain
    dg = graph(src[:E], dst[:E], t[:E], x[:E])
    torch.manual_seed(0)
    mem = TGNMemory(N, D, M, TD).to(DEV).eval()
    enc = GraphAttentionEmbedding(M, 100, D, mem.time_enc).to(DEV).eval()
    hm = HookManager(keys=['k'])
    hm.register('k', RandomNegativeEdgeSamplerHook(low=8227, high=N))
    hm.register('k', RecencyNeighborHook(num_nodes=N, num_nbrs=[k],
                                         seed_nodes_keys=['edge_src', 'edge_dst', 'neg'],
                                         seed_times_keys=['edge_time', 'edge_time', 'neg_time']))
    hm.register('k', DeduplicationHook(seed_nodes_keys=['neg', 'nbr_nids']))
    keep = {}

    def step(batch):
        # non-padded slots of the sampled frontier: one tgm_frontier_compact + gathers (the example's
        # four boolean-mask selections are four select kernels + four count read-backs)
        nbr = batch.nbr_nids[0].flatten()
        idx = compact_frontier(nbr)
        seeds = torch.cat([batch.edge_src, batch.edge_dst, batch.neg])
        ei = torch.stack([batch.global_to_local(seeds.index_select(0, idx // k)),
                          batch.global_to_local(nbr.index_select(0, idx))]).long()
        et = batch.nbr_edge_time[0].flatten().index_select(0, idx)
        ex = batch.nbr_edge_x[0].flatten(0, -2).index_select(0, idx)
        z, lu = mem(batch.unique_nids)
        emb = enc(z, lu, ei, et, ex)
        mem.update_state(batch.edge_src, batch.edge_dst, batch.edge_time, batch.edge_x)
        return z, lu, ei, et, ex, emb

    def epoch():
        hm.reset_state()
        mem.reset_state()
        with hm.activate('k'):
            for i, batch in enumerate(DGDataLoader(dg, batch_size=bs, hook_manager=hm)):
                out = step(batch)
                if i == nb - 1:
                    keep['last'] = out
    torch.set_grad_enabled(False)
    epoch()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    epoch()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / nb * 1e3
    z, lu, ei, et, ex, emb = [v.cpu().numpy() for v in keep['last']]
    p = {k_: v.detach().cpu().numpy() for k_, v in enc.state_dict().items()}
    t_cpu = cpu_s(lambda: graph_attention_embedding(p, 2, z, lu, ei, et, ex))
    want = graph_attention_embedding(p, 2, z, lu, ei, et, ex)
    ms_enc = cuda_ms(lambda: enc(*keep['last'][:5]), iters=20)
    emit('N4 full TGN inference step per batch through DGDataLoader + hooks (wiki-shaped, bs=200 + 200 '
         'negatives, k=10, memory 100, embed 100)', value=ms, unit='ms/batch', higher_is_better=False,
         events_per_s=bs / (ms * 1e-3), embedding_ms=ms_enc, embedding_edges=int(ei.shape[1]),
         embedding_nodes=int(z.shape[0]), embedding_max_abs_err_vs_oracle=float(np.abs(emb - want).max()),
         note='negatives -> recency sampler (ring) -> dedup -> TGNMemory.forward -> '
              'GraphAttentionEmbedding (2 SGEMM + sort + 3 kernels) -> update_state; wall clock, '
              'launch/Python-bound at bs=200; TransformerConv parity is unpinned (third party)',
         cpu_baseline={'value': t_cpu * 1e3, 'unit': 'ms/batch (embedding only)', 'kind': 'port',
                       'cores': os.cpu_count(), 'sample': 'numpy oracle of the last batch'})


# ---- A5 DyGFormer (config 5) ------------------------------------------------------------------------
def row_dygformer():
    from oracle import nn_oracle
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    N, B, L, dN, dE, dT, C, out = 100_000, 200, 32, 128, 16, 100, 50, 172
    m = DyGFormer(dN, dE, dT, C, output_dim=out, patch_size=1, num_layers=2, num_heads=2,
                  max_input_sequence_length=L).to(DEV).eval()
    k = L - 1
    node_x = torch.randn(N, dN, device=DEV)
    src, dst = rng.integers(0, N, B), rng.integers(0, N, B)
    t = rng.integers(10_000, 2_000_000, B)
    nbrs = rng.integers(0, N, (2 * B, k)).astype(np.int32)
    nt = np.sort(np.clip(np.tile(t, 2)[:, None] - rng.integers(1, 9000, (2 * B, k)), 0, None), 1)
    ef = rng.standard_normal((2 * B, k, dE)).astype(np.float32)
    dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    args = (node_x, dv(np.stack([src, dst])), dv(t), dv(nbrs), dv(nt), dv(ef))
    ms = cuda_ms(lambda: m(*args), iters=10)
    p = {k_: v.detach().cpu().numpy() for k_, v in m.state_dict().items()}
    npx = node_x.cpu().numpy()
    t_cpu = cpu_s(lambda: nn_oracle.dygformer_forward(p, 1, 2, 2, npx, np.stack([src, dst]), t, nbrs, nt, ef))
    from oracle import torch_eager as te
    pt = {k_: v.detach() for k_, v in m.state_dict().items()}
    eargs = (node_x, args[1].long(), args[2], args[3].long(), args[4], args[5])
    eager_fwd = lambda: te.dygformer_forward(pt, 1, 2, 2, *eargs)  # noqa: E731
    with torch.no_grad():
        zs, zd = m(*args)
        es, ed = eager_fwd()
        err = float(max((zs - es).abs().max(), (zd - ed).abs().max()))
        ms_eager = cuda_ms(eager_fwd, iters=10)
    E_ = 4 * C
    tok = B * 2 * L
    flops = 2 * (2 * tok * E_ * (3 * E_ + E_ + 8 * E_) + 2 * 2 * B * 2 * (2 * L) ** 2 * (E_ // 2))
    emit('A5 DyGFormer forward (200 edges -> 400 sequences of 32, 4x50 channels, 2 layers)', value=ms,
         unit='ms/call', higher_is_better=False, dense_gflop=flops / 1e9,
         note='frontend kernel (gather, Time2Vec, exact co-occurrence counts) + transformer: token linears on '
              'the tensor cores (tcgen05, fp32-accurate 9xBF16 emulation, fused bias/residual/GELU), '
              'per-head attention products on cuBLAS fp32',
         cpu_baseline={'value': t_cpu * 1e3, 'unit': 'ms/call', 'kind': 'port', 'cores': os.cpu_count(),
                       'sample': 'numpy oracle of the same call'},
         eager_cuda_baseline={'value': ms_eager, 'unit': 'ms/call', 'max_abs_diff_vs_b200_path': err,
                              'sample': 'oracle/torch_eager.py::dygformer_forward on cuda:0 (cuBLAS '
                                        'TF32 off: torch default fp32 matmul), same call, same weights'})


ROWS = {'tgn_step': row_tgn_step, 'store': row_store, 'ring': row_ring, 'twohop': row_twohop, 'forms': row_forms, 'uniform': row_uniform,
        'stream': row_stream_kernels, 'tgat': row_tgat, 'tgn': row_tgn, 'dygformer': row_dygformer}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', default=','.join(ROWS))
    a = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit('bench_rows.py: no CUDA device; the B200 path has no CPU fallback')
    for name in a.rows.split(','):
        try:
            ROWS[name]()
        except Exception as e:  # noqa: BLE001  keep the other rows
            emit(name, error=repr(e))
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
