#!/usr/bin/env python
"""Headline benchmark: sampled-edges/sec of recent-neighbor sampling (k=20) on a synthetic CTDG.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # B200 path (this repo)
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU oracle port, host cores

A *step* is one pass of the hot path over one window of `--window-batches` loader batches
(batch_size 200 stream edges, seeds = both endpoints of every edge, k most recent neighbours
each, D-float edge features gathered): one `tgm_csr_sample_edges` launch.  A *sampled edge* is
one (seed, neighbour-slot) pair returned.

  value  device-resident inputs and outputs, CUDA-event time of exactly K steps, max over ranks
  e2e    the same work through the host-buffer C-ABI call (`tgm_csr_sample_edges_host`): every
         step uploads its slab of stream edges from pinned host memory and downloads the full
         sampled output to pinned host memory
  N>1    the batch stream is time-range sharded, store + adjacency replicated, no data-path
         collective; weak scaling: every rank runs K steps of its shard

Only the cpu_baseline / --impl reference legs touch oracle/ (as the thing being timed there).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'sampled-edges/sec (k=20) on 100M-edge CTDG at 1/2/4/8 B200; HBM GB/s %peak'
try:  # the driver's own wording of the metric, when the file travels with the repo
    METRIC = json.load(open(os.path.join(ROOT, 'BASELINE.json')))['metric']
except Exception:  # noqa: BLE001
    pass
UNIT = 'sampled-edges/s'


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=100)
    p.add_argument('--warmup', type=int, default=5)
    p.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    p.add_argument('--edges', type=int, default=100_000_000)
    p.add_argument('--nodes', type=int, default=1_000_000)
    p.add_argument('--t-max', type=int, default=2000, help='timestamps ~ U{0..t_max-1}; '
                   'nodes*t_max < 2^31 keeps the stream inside the reference parity domain')
    p.add_argument('--dim', type=int, default=16)
    p.add_argument('--k', type=int, default=20)
    p.add_argument('--batch-size', type=int, default=200)
    p.add_argument('--window-batches', type=int, default=5000)
    p.add_argument('--e2e-window-batches', type=int, default=1000)
    p.add_argument('--e2e-steps', type=int, default=0, help='0 = same as --steps')
    p.add_argument('--cpu-sample-edges', type=int, default=9_000_000)
    p.add_argument('--no-cpu-baseline', action='store_true')
    p.add_argument('--no-e2e', action='store_true')
    p.add_argument('--no-numa-bind', action='store_true')
    p.add_argument('--loader-batches', type=int, default=20000,
                   help='loader batches iterated for the public-API (DGDataLoader + hook) number; 0 = skip')
    p.add_argument('--no-colocate', action='store_true')
    p.add_argument('--feature-copy', default='tma', choices=['tma', 'lsu'],
                   help='how the sampler moves feature rows: TMA bulk copies or warp loads/stores')
    p.add_argument('--tma-ctas-per-sm', type=int, default=0)
    p.add_argument('--seed', type=int, default=0)
    return p.parse_args()


def workload_config(a, world):
    return {
        'workload': f'synthetic CTDG {a.edges // 1_000_000}M edges / {a.nodes // 1000}k nodes, '
                    f'D={a.dim} edge features, recent-neighbor sampling k={a.k}, loader '
                    f'batch_size={a.batch_size}, seeds=[edge_src|edge_dst], undirected',
        'edges': a.edges, 'nodes': a.nodes, 't_max': a.t_max, 'edge_x_dim': a.dim, 'k': a.k,
        'batch_size': a.batch_size, 'window_batches': a.window_batches,
        'sharding': f'time-range x{world}, store+adjacency replicated, no collective',
        'l2': 'every step reads a different window; bytes touched per step >> 126 MB L2',
    }


# ---- clocks ------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed regions (NVML)."""
    BAD = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20}
    NOTE = {'sw_power_cap': 0x4}

    def __init__(self, index: int, period: float = 0.02) -> None:
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self._nv = None
            self.error = repr(e)
            return
        self._period = period
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    except Exception:  # noqa: BLE001
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for name, bit in {**self.BAD, **self.NOTE}.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(self._period)

    def start(self):
        self._active.set()

    def pause(self):
        self._active.clear()

    def result(self):
        self._stop.set()
        if self._nv is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'error': self.error}
        return {'sm_mhz': statistics.median(self.samples) if self.samples else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


# ---- NUMA placement ---------------------------------------------------------------------------
def bind_to_gpu_numa_node(index: int):
    """Pin this rank's threads to the CPUs of the NUMA node its GPU hangs off, BEFORE any host
    buffer is allocated (pinned pages then come from that node: first-touch policy).  The e2e leg
    moves ~0.6 GB per step between host DRAM and the GPU; with every rank's buffers on one socket
    the ranks share that socket's memory and inter-socket links.  Returns what was done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(':')[0]) == 8:  # NVML pads the PCI domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open(f'/sys/bus/pci/devices/{bus}/numa_node').read())
        if node < 0:
            return {'node': None, 'why': 'no NUMA affinity reported for the GPU'}
        cpus = set()
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {'node': node, 'why': 'node CPUs not in this process\'s affinity mask'}
        os.sched_setaffinity(0, cpus)
        return {'node': node, 'cpus': len(cpus)}
    except Exception as e:  # noqa: BLE001  placement is an optimisation, never a failure
        return {'node': None, 'why': repr(e)}


# ---- CPU baseline (the only place oracle/ is executed, as the thing timed) -----------------------
def numpy_prefix_stream(a, n_edges):
    import numpy as np
    rng = np.random.default_rng(a.seed)
    src = rng.integers(0, a.nodes, n_edges).astype(np.int32)
    dst = rng.integers(0, a.nodes, n_edges).astype(np.int32)
    t_hi = max(1, int(a.t_max * n_edges / a.edges))  # a prefix of the stream covers early times
    t = np.sort(rng.integers(0, t_hi, n_edges)).astype(np.int64)
    x = rng.standard_normal((n_edges, a.dim)).astype(np.float32) if a.dim else None
    return src, dst, t, x


def time_cpu_port(a, stream, n_edges, warm_edges=0):
    """C port of the reference ring sampler (oracle/recency_ring.c) over a prefix of the stream,
    loader batch by loader batch: query both endpoints, then push the batch."""
    from oracle.c_oracle import CRing
    src, dst, t, x = stream
    ring = CRing(a.nodes, [a.k], a.dim)
    if warm_edges:
        ring.run_stream(src, dst, t, x, 0, warm_edges, a.batch_size)
    t0 = time.perf_counter()
    slots, _, _ = ring.run_stream(src, dst, t, x, warm_edges, n_edges, a.batch_size)
    dt = time.perf_counter() - t0
    return slots / dt, slots, dt


def run_reference(a):
    """`--impl reference`: the reference's algorithm on the host cores.  The reference is pure
    Python and cannot travel to the GPU box, so this times the C port of its state machine
    (a far faster stand-in than the reference's eager-PyTorch implementation: SURVEY.md section 6
    measured the real one at ~4e5 sampled-edges/s at this geometry)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle.c_oracle import CRing
    step_edges = 250 * a.batch_size  # bounded sample per step
    need = (a.steps + a.warmup) * step_edges
    src, dst, t, x = numpy_prefix_stream(a, need)
    ring = CRing(a.nodes, [a.k], a.dim)
    at = 0
    for _ in range(a.warmup):
        ring.run_stream(src, dst, t, x, at, at + step_edges, a.batch_size)
        at += step_edges
    t0 = time.perf_counter()
    slots = 0
    for _ in range(a.steps):
        s, _, _ = ring.run_stream(src, dst, t, x, at, at + step_edges, a.batch_size)
        slots += s
        at += step_edges
    dt = time.perf_counter() - t0
    value = slots / dt
    sample = (f'{a.steps} steps x 250 loader batches (bs={a.batch_size}) from the head of the '
              f'stream after {a.warmup} warm-up steps; C port of the ring sampler, 1 thread '
              f'(the state machine is sequential across batches)')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': dt / a.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32/int64 '
        'ids+times, f32 feature copy', 'data': 'synthetic',
        'config': workload_config(a, 1),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                         'sample': sample, 'host_cores': os.cpu_count()},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


# ---- B200 arm -----------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist

    from tgm_b200 import RecencyCSR, _cabi
    from tgm_b200.core.storage import DeviceCOOStorage
    _cabi.check(_cabi.lib.tgm_set_option(b'csr_feature_copy', int(a.feature_copy == 'tma')))
    _cabi.check(_cabi.lib.tgm_set_option(b'csr_tma_ctas_per_sm', a.tma_ctas_per_sm))

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the B200 path has no CPU fallback')
    numa = bind_to_gpu_numa_node(local) if not a.no_numa_bind else {'node': None, 'why': 'disabled'}
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION/INFO; keep stdout to the
        # one JSON line the driver parses
        os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=dev)

    E, N, D, k, bs = a.edges, a.nodes, a.dim, a.k, a.batch_size
    # synthetic stream, identical on every rank (same generator seed, same device type)
    gen = torch.Generator(device=dev).manual_seed(a.seed)
    src = torch.randint(0, N, (E,), generator=gen, device=dev, dtype=torch.int32)
    dst = torch.randint(0, N, (E,), generator=gen, device=dev, dtype=torch.int32)
    t = torch.sort(torch.randint(0, a.t_max, (E,), generator=gen, device=dev))[0]
    x = torch.randn((E, D), generator=gen, device=dev) if D else None
    t_build = time.perf_counter()
    store = DeviceCOOStorage.from_device_tensors(src, dst, t, x, N)
    csr = RecencyCSR(store, bs, colocate_x=not a.no_colocate)
    torch.cuda.synchronize(dev)
    t_build = time.perf_counter() - t_build

    # this rank's time-range shard of the batch stream (tgm_b200/parallel.py)
    from tgm_b200.parallel import shard_batches
    shard = shard_batches(E, bs, rank, world)
    b_lo, b_hi = shard.batch_lo, shard.batch_hi
    W = min(a.window_batches, b_hi - b_lo)
    nwin = max(1, (b_hi - b_lo) // W)

    def window(i, wb=W, count=nwin):
        j = i % count
        lo = (b_lo + j * wb) * bs
        return lo, min(lo + wb * bs, E, b_hi * bs)

    max_edges = W * bs
    out = (torch.empty((2 * max_edges, k), dtype=torch.int32, device=dev),
           torch.empty((2 * max_edges, k), dtype=torch.int64, device=dev),
           torch.empty((2 * max_edges, k, D), dtype=torch.float32, device=dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(i):
        lo, hi = window(i)
        n = 2 * (hi - lo)
        csr.sample_edges(lo, hi, k, k, out=(out[0][:n], out[1][:n], out[2][:n]))
        return n * k

    clocks = ClockSampler(local)
    for i in range(max(a.warmup, 3)):  # never fewer than 3 untimed steps
        step(i)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(a.steps)]
    clocks.start()
    slots = 0
    for i in range(a.steps):
        ev[i][0].record()
        slots += step(max(a.warmup, 3) + i)
        ev[i][1].record()
    barrier()
    clocks.pause()
    kernel_ms = [s.elapsed_time(e) for s, e in ev]
    total_ms = ev[0][0].elapsed_time(ev[-1][1])

    def reduce_max(v):
        if world == 1:
            return v
        tt = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def reduce_sum(v):
        if world == 1:
            return v
        tt = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        return float(tt.item())

    total_ms_max = reduce_max(total_ms)
    slots_all = reduce_sum(float(slots))
    value = slots_all / (total_ms_max * 1e-3)

    # roofline of the dominant (only) kernel: algorithmic bytes per launch / mean launch duration
    slots_per_launch = slots / a.steps
    bytes_per_slot = 2 * (12 + 4 * D) + 20.0 / k
    algo_bytes = slots_per_launch * bytes_per_slot
    launch_s = reduce_max(statistics.mean(kernel_ms)) * 1e-3  # the slowest rank's launches
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    achieved = algo_bytes / launch_s / 1e9
    traffic = None  # DRAM bytes per launch from the committed ncu capture of this kernel
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
        if a.feature_copy == 'tma' and D == 16 and k == 20:
            traffic = tr['dram_bytes_per_sampled_edge'] * slots_per_launch
    except Exception:  # noqa: BLE001
        pass
    roofline = {'bound': 'hbm', 'kernel': 'csr_sample_tma_kernel<true>' if a.feature_copy == 'tma' else
                'csr_sample_edges_fast_kernel', 'achieved': achieved,
                'peak': peak, 'peak_source': 'measured' if 'hbm_gbs' in peaks else 'fallback',
                'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                'traffic_source': 'profiles/traffic.json (ncu dram__bytes_read+write per sampled edge '
                                  'x sampled edges per launch)' if traffic else None,
                'algorithmic_bytes_per_launch': algo_bytes,
                'bytes_per_sampled_edge': bytes_per_slot, 'launch_ms': launch_s * 1e3}

    # ---- e2e: host buffers in, host buffers out, through the C ABI --------------------------
    e2e = None
    if not a.no_e2e:
        We = min(a.e2e_window_batches, b_hi - b_lo)
        n_host_win = min(8, max(1, (b_hi - b_lo) // We))
        ke = a.e2e_steps or a.steps
        me = We * bs
        host_in = []
        for j in range(n_host_win):
            lo, hi = window(j, We, n_host_win)
            host_in.append((lo, hi, tuple(
                None if v is None else v[lo:hi].cpu().pin_memory() for v in (src, dst, t, x))))
        nslot = 2
        host_out = [(torch.empty((2 * me, k), dtype=torch.int32).pin_memory(),
                     torch.empty((2 * me, k), dtype=torch.int64).pin_memory(),
                     torch.empty((2 * me, k, D), dtype=torch.float32).pin_memory())
                    for _ in range(nslot)]
        streams = [torch.cuda.Stream(dev) for _ in range(nslot)]

        def e2e_step(i):
            lo, hi, hin = host_in[i % n_host_win]
            sl = i % nslot
            streams[sl].synchronize()  # the previous result in this slot has been consumed
            csr.sample_edges_host(lo, hi, k, k, hin, host_out[sl], slot=sl,
                                  stream=streams[sl].cuda_stream)
            return 2 * (hi - lo) * k, hin

        for i in range(max(3, a.warmup)):
            e2e_step(i)
        barrier()
        clocks.start()
        t0 = time.perf_counter()
        eslots = 0
        for i in range(ke):
            s, hin = e2e_step(i)
            eslots += s
        for s_ in streams:
            s_.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        clocks.pause()
        h2d = sum(v.numel() * v.element_size() for v in hin if v is not None)
        d2h = 2 * (host_in[0][1] - host_in[0][0]) * k * (4 + 8 + 4 * D)
        # spot-check: the host result equals the device-resident path on the same window
        lo, hi, hin = host_in[(ke - 1) % n_host_win]
        sl = (ke - 1) % nslot
        ref = csr.sample_edges(lo, hi, k, k)
        n = 2 * (hi - lo)
        assert torch.equal(host_out[sl][0][:n], ref[0].cpu()), 'e2e ids differ from device path'
        assert torch.equal(host_out[sl][1][:n], ref[1].cpu())
        assert torch.equal(host_out[sl][2][:n], ref[2].cpu())
        dt_max = reduce_max(dt)
        e2e = {'value': reduce_sum(float(eslots)) / dt_max, 'unit': UNIT,
               'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': ke,
               'ms_per_step': dt_max / ke * 1e3, 'window_batches': We,
               'api': 'tgm_csr_sample_edges_host (pinned host slab in, pinned host result out, '
                      '2 streams)'}
        del host_out, host_in

    # ---- the call a TGM user makes: for batch in DGDataLoader(dg, 200, hook_manager=hm) ---------
    loader_api = None
    if a.loader_batches and not a.no_e2e:
        from tgm_b200 import DGDataLoader, DGraph, HookManager, RecencyNeighborHook
        from tgm_b200.core.storage import DGSliceTracker
        from tgm_b200.core.timedelta import TimeDeltaDG
        nbl = min(a.loader_batches, b_hi - b_lo)
        # every rank walks the head of the stream (loader.py:137-139 iterates absolute event indices
        # from 0, so an index-sliced view cannot start mid-stream); this key measures the Python API
        sl = DGSliceTracker(end_idx=min(nbl * bs, E))
        dg = DGraph._from_storage(store, TimeDeltaDG('r'), dev, sl)
        hm = HookManager(keys=['bench'])
        hm.register('bench', RecencyNeighborHook(
            num_nodes=N, num_nbrs=[k], seed_nodes_keys=['edge_src', 'edge_dst'],
            seed_times_keys=['edge_time', 'edge_time'], window_batches=min(5000, nbl)))
        with hm.activate('bench'):
            for pass_ in range(2):  # first pass warms the adjacency cache and the allocator
                hm.reset_state()
                barrier()
                t0 = time.perf_counter()
                got = 0
                for batch in DGDataLoader(dg, batch_size=bs, hook_manager=hm):
                    got += batch.nbr_nids[0].numel()
                barrier()
                dt = time.perf_counter() - t0
        dt_max = reduce_max(dt)
        loader_api = {'value': reduce_sum(float(got)) / dt_max, 'unit': UNIT, 'batches': nbl,
                      'us_per_batch': dt_max / nbl * 1e6,
                      'api': 'DGDataLoader(batch_size=200) + HookManager + RecencyNeighborHook('
                             'window_batches=5000): outputs stay on the device, one DGBatch per '
                             'iteration (Python-bound)'}

    # ---- CPU baseline on this host (rank 0, N=1 only) ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        n_cpu = min(a.cpu_sample_edges, E)
        warm = n_cpu // 3
        stream = tuple(None if v is None else v[:n_cpu].cpu().numpy() for v in (src, dst, t, x))
        v, s, dt = time_cpu_port(a, stream, n_cpu, warm)
        # the numpy restatement follows the reference's eager tensor-op structure (gathers, masks,
        # argsort) more closely than the C port does: reported beside it, on fewer batches
        from oracle.recency_oracle import RingOracle
        nb_np = 300
        ring_np = RingOracle(N, [k], D)
        lo_np = warm
        t0 = time.perf_counter()
        for b in range(nb_np):
            lo_b, hi_b = lo_np + b * bs, lo_np + (b + 1) * bs
            sd = np.concatenate([stream[0][lo_b:hi_b], stream[1][lo_b:hi_b]])
            tq_ = np.concatenate([stream[2][lo_b:hi_b]] * 2)
            ring_np.hook_call(sd, tq_, stream[0][lo_b:hi_b], stream[1][lo_b:hi_b],
                              stream[2][lo_b:hi_b], None if D == 0 else stream[3][lo_b:hi_b])
        np_rate = nb_np * 2 * bs * k / (time.perf_counter() - t0)
        cpu = {'value': v, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'host_cores': os.cpu_count(),
               'numpy_port_value': np_rate,
               'sample': f'edges [{warm}, {n_cpu}) of the same stream ({s} sampled edges, '
                         f'{dt:.1f} s) after pushing the first {warm}; C port of the '
                         f'reference ring sampler (oracle/recency_ring.c), batch by batch'}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': total_ms_max / a.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'int32/int64 ids+times, f32 feature copy', 'data': 'synthetic',
            'config': workload_config(a, world), 'gpu_launches': a.steps,
            'stream_edges_per_s': value / (2 * k),
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'loader_api': loader_api,
            'clocks': clocks.result(), 'numa': numa,
            'build_s': t_build,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)


if __name__ == '__main__':
    main()
