#!/usr/bin/env python
"""Headline benchmark: sampled-edges/sec of recent-neighbor sampling (k=20) on a synthetic CTDG.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # B200 path (this repo)
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU oracle port, host cores

A *step* is one pass of the hot path over one window of `--window-batches` loader batches
(batch_size 200 stream edges, seeds = both endpoints of every edge, k most recent neighbours
each, D-float edge features gathered): one `tgm_csr_sample_edges` launch.  A *sampled edge* is
one (seed, neighbour-slot) pair returned.

  value  device-resident inputs and outputs, CUDA-event time of exactly K steps, max over ranks;
         the timed windows come from the steady state of the stream (every ring full)
  e2e    the same sampling through the host-buffer C-ABI call `tgm_csr_sample_edges_host_ids`:
         every step uploads its slab of stream edges (src, dst, t) from pinned host memory, the
         kernel reads its seeds from that slab, and (nid, t, eid) -- 16 bytes per sampled edge,
         what a host that owns the feature table lacks -- come back to pinned host memory;
         `e2e.variants` holds the full-row form (12 + 4D bytes per sampled edge back) and the
         fused sample + masked-mean form (4D bytes per seed back)
  full_pass  one pass over the whole stream including the build of the adjacency
         (`value_incl_build`)
  N>1    the batch stream is time-range sharded, store + adjacency replicated, no data-path
         collective for sampling; weak scaling: every rank runs K steps of
         its shard.  `collective` times the one exchange the path has: the TGN node-memory join
         at shard boundaries (rows each shard touched, all-gathered over NCCL)

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
    p.add_argument('--cpu-sample-edges', type=int, default=4_000_000,
                   help='stream edges the CPU baseline is timed on (after a steady-state warm-up)')
    p.add_argument('--e2e-full-steps', type=int, default=20)
    p.add_argument('--no-cpu-baseline', action='store_true')
    p.add_argument('--no-e2e', action='store_true')
    p.add_argument('--no-numa-bind', action='store_true')
    p.add_argument('--no-collective', action='store_true')
    p.add_argument('--join-edges', type=int, default=100_000,
                   help='stream edges per shard whose endpoints count as touched in the memory '
                        'join (500 loader batches of 200, as bench_tgn_shard.py runs between joins)')
    p.add_argument('--loader-batches', type=int, default=20000,
                   help='loader batches iterated for the public-API (DGDataLoader + hook) number; 0 = skip')
    p.add_argument('--no-colocate', action='store_true')
    p.add_argument('--feature-copy', default='tma', choices=['tma', 'lsu'],
                   help='how the sampler moves feature rows: TMA bulk copies or warp loads/stores')
    p.add_argument('--tma-ctas-per-sm', type=int, default=0)
    p.add_argument('--seed', type=int, default=0)
    return p.parse_args()


def steady_start_edge(a):
    """First stream edge of the steady state: from here on a node has seen on average >= 2.5 k
    entries, so practically every ring holds k neighbours and no sampled slot is padding."""
    e = int(1.25 * a.k * a.nodes)
    e -= e % a.batch_size
    return min(e, a.edges // 2 // a.batch_size * a.batch_size)


def workload_config(a, world):
    return {
        'workload': f'synthetic CTDG {a.edges // 1_000_000}M edges / {a.nodes // 1000}k nodes, '
                    f'D={a.dim} edge features, recent-neighbor sampling k={a.k}, loader '
                    f'batch_size={a.batch_size}, seeds=[edge_src|edge_dst], undirected',
        'edges': a.edges, 'nodes': a.nodes, 't_max': a.t_max, 'edge_x_dim': a.dim, 'k': a.k,
        'batch_size': a.batch_size, 'window_batches': a.window_batches,
        'timed_region': f'steady state: windows drawn from stream edges >= {steady_start_edge(a)} '
                        f'(mean adjacency length >= 2.5 k, every ring full)',
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
def numpy_stream(a, e_lo, e_hi):
    """Edges [e_lo, e_hi) of a stream with the bench workload's statistics, generated on the host
    (the reference arm runs without touching the GPU)."""
    rng = np.random.default_rng(a.seed)
    n = e_hi - e_lo
    src = rng.integers(0, a.nodes, n).astype(np.int32)
    dst = rng.integers(0, a.nodes, n).astype(np.int32)
    t_lo, t_hi = int(a.t_max * e_lo / a.edges), max(1, int(a.t_max * e_hi / a.edges))
    t = np.sort(rng.integers(t_lo, max(t_hi, t_lo + 1), n)).astype(np.int64)
    x = rng.standard_normal((n, a.dim), dtype=np.float32) if a.dim else None
    return src, dst, t, x


def cpu_port_steady(a, stream, warm_edges, step_edges, steps, warmup):
    """C port of the reference ring sampler (oracle/recency_ring.c), loader batch by loader batch:
    query both endpoints, then push the batch.  The first `warm_edges` edges of `stream` are pushed
    without querying so that the timed batches see full rings (steady state, like the GPU arm)."""
    from oracle.c_oracle import CRing
    src, dst, t, x = stream
    ring = CRing(a.nodes, [a.k], a.dim)
    t0 = time.perf_counter()
    ring.push_stream(src, dst, t, x, 0, warm_edges, a.batch_size)
    warm_s = time.perf_counter() - t0
    at = warm_edges
    for _ in range(warmup):
        ring.run_stream(src, dst, t, x, at, at + step_edges, a.batch_size, checksum=False)
        at += step_edges
    t0 = time.perf_counter()
    slots = 0
    for _ in range(steps):
        sl, _, _ = ring.run_stream(src, dst, t, x, at, at + step_edges, a.batch_size,
                                   checksum=False)
        slots += sl
        at += step_edges
    dt = time.perf_counter() - t0
    return slots / dt, slots, dt, warm_s


def run_reference(a):
    """`--impl reference`: the reference's algorithm on the host cores.  The reference is pure
    Python and cannot travel to the GPU box, so this times the C port of its state machine
    (a far faster stand-in than the reference's eager-PyTorch implementation: SURVEY.md section 6
    measured the real one at ~4e5 sampled-edges/s at this geometry).  One thread: the state
    machine is sequential across loader batches and one batch is 400 seeds."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    step_edges = 250 * a.batch_size  # bounded sample per step
    # the ring of a node holds its last k entries: pushing the `warm` edges that precede the timed
    # region reproduces the steady state (a node sees on average 2 * warm / nodes >= 2.5 k entries)
    warm = steady_start_edge(a)
    need = warm + (a.steps + a.warmup) * step_edges
    stream = numpy_stream(a, 0, need)
    value, slots, dt, warm_s = cpu_port_steady(a, stream, warm, step_edges, a.steps, a.warmup)
    sample = (f'{a.steps} steps x 250 loader batches (bs={a.batch_size}) starting at stream edge '
              f'{warm + a.warmup * step_edges} (steady state: the preceding {warm} edges were pushed '
              f'first, {warm_s:.0f} s untimed); C port of the ring sampler, 1 thread (the state '
              f'machine is sequential across batches)')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': dt / a.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32/int64 '
        'ids+times, f32 feature copy', 'data': 'synthetic',
        'config': workload_config(a, a.gpus),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                         'sample': sample, 'host_cores': os.cpu_count()},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def time_eager_cuda(a, csr, src, dst, t, x, dev, batches=300):
    """The reference hook's tensor-op sequence (oracle/torch_eager.py::TorchRing: the O(N*B) `.min()`
    scan + host sync per call, ~25 eager ops per query, argsort + ~15 ops per push;
    recency.py:239-399) run on the B200 itself, batch by batch, in the steady state: its ring
    buffers are loaded with the exact state after the first `steady_start_edge` stream edges
    (tgm_csr_export_ring), the first batch's answer is checked against the CUDA sampler."""
    import ctypes

    import torch

    from oracle.torch_eager import TorchRing
    from tgm_b200 import _cabi
    N, D, k, bs = a.nodes, a.dim, a.k, a.batch_size
    e0 = steady_start_edge(a)
    h = ctypes.c_void_p()
    _cabi.check(_cabi.lib.tgm_recency_create(ctypes.byref(h), N, k, D, dev.index))
    try:
        _cabi.check(_cabi.lib.tgm_csr_export_ring(csr.handle, e0, h, _cabi.current_stream(dev)))
        p = [ctypes.c_void_p() for _ in range(4)]
        _cabi.check(_cabi.lib.tgm_recency_state(h, *[ctypes.byref(q) for q in p]))
        ring = TorchRing(N, [k], D, device=dev)
        ring.ids.copy_(_cabi.device_view(p[0].value, (N, k), torch.int32, dev))
        ring.times.copy_(_cabi.device_view(p[1].value, (N, k), torch.int64, dev))
        if D:
            ring.feats.copy_(_cabi.device_view(p[2].value, (N, k, D), torch.float32, dev))
        ring.write_pos.copy_(_cabi.device_view(p[3].value, (N,), torch.int32, dev))
        torch.cuda.synchronize(dev)
    finally:
        _cabi.lib.tgm_recency_destroy(h)

    def batch(i):
        lo, hi = e0 + i * bs, e0 + (i + 1) * bs
        seeds = torch.cat([src[lo:hi], dst[lo:hi]])
        tq = torch.cat([t[lo:hi], t[lo:hi]])
        return ring.hook_call(seeds, tq, src[lo:hi], dst[lo:hi], t[lo:hi],
                              None if x is None else x[lo:hi])
    got = batch(0)[0]
    want = csr.sample_edges(e0, e0 + bs, k, k)
    same = all(torch.equal(u, v) for u, v in zip(got[2:], want))
    for i in range(1, 20):
        batch(i)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(20, 20 + batches):
        batch(i)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    return {'value': batches * 2 * bs * k / dt, 'unit': UNIT, 'us_per_batch': dt / batches * 1e6,
            'kind': 'eager-torch restatement of the reference hook (oracle/torch_eager.py) on the '
                    'same B200: what the reference\'s device=\'cuda\' mode launches',
            'first_batch_equals_cuda_sampler': bool(same),
            'sample': f'{batches} loader batches from stream edge {e0 + 20 * bs} on, hook only (no '
                      f'O(E) materialize), steady-state rings'}


# ---- B200 arm -----------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist

    from tgm_b200 import RecencyCSR, _cabi
    from tgm_b200.core.storage import DeviceCOOStorage
    _cabi.check(_cabi.lib.tgm_set_option(b'csr_feature_copy', int(a.feature_copy == 'tma')))
    _cabi.check(_cabi.lib.tgm_set_option(b'csr_tma_ctas_per_sm', a.tma_ctas_per_sm))
    if os.environ.get('TGM_B200_TRACE'):
        _cabi.check(_cabi.lib.tgm_set_option(b'trace', 1))

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the B200 path has no CPU fallback')
    # stdout carries exactly ONE line, the JSON the driver parses: anything native code prints
    # there (NCCL's version banner does, whatever NCCL_DEBUG_FILE says) is sent to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    numa = bind_to_gpu_numa_node(local) if not a.no_numa_bind else {'node': None, 'why': 'disabled'}
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # stdout carries the one JSON line the driver parses: NCCL's own log (whatever level the
        # environment asks for; INIT lines by default so the communicator size is on record) goes
        # to stderr
        # never lowered, never silenced: a quieter preset (this pool's boxes export
        # NCCL_DEBUG=VERSION) is raised to INFO / INIT so the communicator size is on record
        if os.environ.get('NCCL_DEBUG', '').upper() not in ('INFO', 'TRACE'):
            os.environ['NCCL_DEBUG'] = 'INFO'
            os.environ.setdefault('NCCL_DEBUG_SUBSYS', 'INIT')
        # (no NCCL_DEBUG_FILE: NCCL logs to fd 1, which now IS stderr; reopening /dev/stderr by
        # name would truncate it when stderr is a regular file)
        dist.init_process_group('nccl', device_id=dev)

    E, N, D, k, bs = a.edges, a.nodes, a.dim, a.k, a.batch_size
    # synthetic stream, identical on every rank (same generator seed, same device type)
    gen = torch.Generator(device=dev).manual_seed(a.seed)
    src = torch.randint(0, N, (E,), generator=gen, device=dev, dtype=torch.int32)
    dst = torch.randint(0, N, (E,), generator=gen, device=dev, dtype=torch.int32)
    t = torch.sort(torch.randint(0, a.t_max, (E,), generator=gen, device=dev))[0]
    x = torch.randn((E, D), generator=gen, device=dev) if D else None
    # process start-up is not build time: one throw-away build of a 100k-edge prefix loads the
    # build kernels (CUDA loads a kernel's code at its first launch) and warms the allocator
    warm_n = min(E, 100_000)
    _w = DeviceCOOStorage.from_device_tensors(src[:warm_n].clone(), dst[:warm_n].clone(),
                                              t[:warm_n].clone(),
                                              None if x is None else x[:warm_n].clone(), N)
    del _w
    _w = RecencyCSR(DeviceCOOStorage.from_device_tensors(
        src[:warm_n].clone(), dst[:warm_n].clone(), t[:warm_n].clone(),
        None if x is None else x[:warm_n].clone(), N), bs, colocate_x=not a.no_colocate)
    del _w
    torch.cuda.synchronize(dev)
    t_build = time.perf_counter()
    store = DeviceCOOStorage.from_device_tensors(src, dst, t, x, N)
    csr = RecencyCSR(store, bs, colocate_x=not a.no_colocate)
    torch.cuda.synchronize(dev)
    t_build = time.perf_counter() - t_build

    # this rank's time-range shard (tgm_b200/parallel.py) of the steady-state part of the batch
    # stream; `full` = its shard of the whole stream (the full pass below)
    from tgm_b200.parallel import shard_batches
    nb_total = (E + bs - 1) // bs
    sb0 = steady_start_edge(a) // bs
    steady = shard_batches((nb_total - sb0) * bs, bs, rank, world)
    b_lo, b_hi = sb0 + steady.batch_lo, sb0 + steady.batch_hi
    full = shard_batches(E, bs, rank, world)
    W = min(a.window_batches, b_hi - b_lo)
    nwin = max(1, (b_hi - b_lo) // W)

    def window(i, wb=W, count=nwin, base=b_lo, end=b_hi):
        j = i % count
        lo = (base + j * wb) * bs
        return lo, min(lo + wb * bs, E, end * bs)

    max_edges = W * bs
    out = (torch.empty((2 * max_edges, k), dtype=torch.int32, device=dev),
           torch.empty((2 * max_edges, k), dtype=torch.int64, device=dev),
           torch.empty((2 * max_edges, k, D), dtype=torch.float32, device=dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def sample_into(lo, hi):
        n = 2 * (hi - lo)
        csr.sample_edges(lo, hi, k, k, out=(out[0][:n], out[1][:n], out[2][:n]))
        return n * k

    def step(i):
        return sample_into(*window(i))

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

    clocks = ClockSampler(local)
    nwarm = max(a.warmup, 3)  # never fewer than 3 untimed steps
    for i in range(nwarm):
        step(i)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(a.steps)]
    clocks.start()
    slots = 0
    for i in range(a.steps):
        ev[i][0].record()
        slots += step(nwarm + i)
        ev[i][1].record()
    barrier()
    clocks.pause()
    kernel_ms = [s.elapsed_time(e) for s, e in ev]
    total_ms = ev[0][0].elapsed_time(ev[-1][1])
    total_ms_max = reduce_max(total_ms)
    slots_all = reduce_sum(float(slots))
    value = slots_all / (total_ms_max * 1e-3)
    # how much of the timed output was padding (steady state: ~0)
    lo_s, hi_s = window(nwarm)
    sample_into(lo_s, hi_s)
    pad_frac = float((out[0][:2 * (hi_s - lo_s)] < 0).float().mean().item())

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
                'bytes_per_sampled_edge': bytes_per_slot, 'launch_ms': launch_s * 1e3,
                'padded_slot_fraction': pad_frac}

    # ---- one pass over the WHOLE stream, build included -------------------------------------
    fW = min(a.window_batches, full.batch_hi - full.batch_lo)
    fwins = [((full.batch_lo + j) * bs, min((full.batch_lo + j + fW) * bs, E, full.batch_hi * bs))
             for j in range(0, full.batch_hi - full.batch_lo, fW)]
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    fslots = sum(sample_into(lo, hi) for lo, hi in fwins)
    f1.record()
    barrier()
    pass_s = f0.elapsed_time(f1) * 1e-3
    build_max, pass_max = reduce_max(t_build), reduce_max(pass_s)
    full_pass = {
        'value_incl_build': reduce_sum(float(fslots)) / reduce_max(t_build + pass_s), 'unit': UNIT,
        'build_s': build_max, 'sample_s': pass_max, 'windows': len(fwins),
        'what': 'every loader batch of the stream sampled once (this rank\'s shard; under-filled '
                'head included), plus the one-off build of store handle + adjacency + anchors + '
                'colocated feature rows from the device-resident edge arrays (timed after a '
                'throw-away 100k-edge build that loads the build kernels)'}

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
        streams = [torch.cuda.Stream(dev) for _ in range(nslot)]
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()  # noqa: E731
        cells = (2 * me, k)

        def run_variant(name, nsteps):
            """K steps of one host-buffer form, double-buffered over two streams; returns
            (sampled edges, seconds, h2d bytes, d2h bytes, spot-check)."""
            if name == 'ids':
                outs = [(pin(cells, torch.int32), pin(cells, torch.int64), pin(cells, torch.int32))
                        for _ in range(nslot)]
                call = csr.sample_edges_host_ids
                nin = 3
            elif name == 'mean':
                outs = [(None, None, pin((2 * me, D), torch.float32)) for _ in range(nslot)]
                call = csr.sample_edges_host_mean
                nin = 3
            else:
                outs = [(pin(cells, torch.int32), pin(cells, torch.int64),
                         pin(cells + (D,), torch.float32)) for _ in range(nslot)]
                call = csr.sample_edges_host
                nin = 4

            def one(i):
                lo, hi, hin = host_in[i % n_host_win]
                sl = i % nslot
                streams[sl].synchronize()  # the previous result in this slot has been consumed
                call(lo, hi, k, k, hin[:nin], outs[sl], slot=sl, stream=streams[sl].cuda_stream)
                return 2 * (hi - lo) * k

            for i in range(3):
                one(i)
            for s_ in streams:
                s_.synchronize()
            barrier()
            clocks.start()
            t0 = time.perf_counter()
            got = 0
            for i in range(nsteps):
                got += one(i)
            for s_ in streams:
                s_.synchronize()
            barrier()
            dt = time.perf_counter() - t0
            clocks.pause()
            lo, hi, hin = host_in[(nsteps - 1) % n_host_win]
            res = outs[(nsteps - 1) % nslot]
            n = 2 * (hi - lo)
            h2d = sum(v.numel() * v.element_size() for v in hin[:nin] if v is not None)
            d2h = sum(v[:n].numel() * v.element_size() for v in res if v is not None)
            # spot-check: the host result equals the device-resident path on the same window
            ref = csr.sample_edges(lo, hi, k, k)
            if name == 'mean':
                want = torch.empty((n, D), dtype=torch.float32, device=dev)
                _cabi.check(_cabi.lib.tgm_masked_mean(ref[2].data_ptr(), ref[0].data_ptr(), n, k, D,
                                                      want.data_ptr(), _cabi.current_stream(dev)))
                assert torch.equal(res[2][:n], want.cpu()), 'e2e fused mean differs'
            else:
                assert torch.equal(res[0][:n], ref[0].cpu()), 'e2e ids differ from device path'
                assert torch.equal(res[1][:n], ref[1].cpu())
                if name == 'ids':  # the host gathers the rows it owns: edge_x[eid]
                    rows = res[2][:4096].to(dev).long()
                    assert torch.equal(x[rows.clamp(min=0)] * (rows >= 0)[..., None],
                                       ref[2][:4096]), 'edge_x[eid] differs from nbr_edge_x'
                else:
                    assert torch.equal(res[2][:n], ref[2].cpu())
            dt_max = reduce_max(dt)
            return {'value': reduce_sum(float(got)) / dt_max, 'unit': UNIT, 'steps': nsteps,
                    'ms_per_step': dt_max / nsteps * 1e3, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h}

        # what the host link itself delivers with every rank copying at once (pinned, one stream)
        lk = pin((64 << 20,), torch.uint8)
        lkd = torch.empty_like(lk, device=dev)
        link = {}
        for name, (dst_, src_) in (('d2h', (lk, lkd)), ('h2d', (lkd, lk))):
            dst_.copy_(src_, non_blocking=True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(8):
                dst_.copy_(src_, non_blocking=True)
            barrier()
            link[name + '_gbs_per_rank'] = 8 * lk.numel() / reduce_max(time.perf_counter() - t0) / 1e9
        del lk, lkd

        e2e = run_variant('ids', ke)
        e2e.update({
            'window_batches': We,
            'api': 'tgm_csr_sample_edges_host_ids (pinned host slab [src|dst|t] in, kernel reads '
                   'its seeds from that slab, pinned (nid, t, eid) out, 2 streams); the host owns '
                   'edge_x and gathers nbr_edge_x = edge_x[eid] itself when it needs rows',
            'link': link,
            'variants': {
                'full_rows': dict(run_variant('rows', min(ke, a.e2e_full_steps)),
                                  api='tgm_csr_sample_edges_host: slab incl. edge_x in, '
                                      '(nid, t, nbr_edge_x) out'),
                'fused_mean': dict(run_variant('mean', ke),
                                   api='tgm_csr_sample_edges_host_mean: slab in, masked mean of '
                                       'the sampled rows (S, D) out') if D and D % 4 == 0 else None,
            }})
        del host_in

    # ---- the call a TGM user makes: for batch in DGDataLoader(dg, 200, hook_manager=hm) ---------
    loader_api = None
    if a.loader_batches and not a.no_e2e:
        from tgm_b200 import (DGDataLoader, DGraph, HookManager, RandomNegativeEdgeSamplerHook,
                              RecencyNeighborHook)
        from tgm_b200.core.storage import DGSliceTracker
        from tgm_b200.core.timedelta import TimeDeltaDG
        nbl = min(a.loader_batches, nb_total)
        # every rank walks the head of the stream (loader.py:137-139 iterates absolute event indices
        # from 0, so an index-sliced view cannot start mid-stream); this key measures the Python API
        sl = DGSliceTracker(end_idx=min(nbl * bs, E))
        dg = DGraph._from_storage(store, TimeDeltaDG('r'), dev, sl)

        def loader_run(with_neg):
            hm = HookManager(keys=['bench'])
            keys_n, keys_t = ['edge_src', 'edge_dst'], ['edge_time', 'edge_time']
            if with_neg:  # the link-prediction recipe (tgm/hooks/recipe.py:51-79)
                hm.register('bench', RandomNegativeEdgeSamplerHook(low=0, high=N))
                keys_n, keys_t = keys_n + ['neg'], keys_t + ['neg_time']
            hm.register('bench', RecencyNeighborHook(
                num_nodes=N, num_nbrs=[k], seed_nodes_keys=keys_n, seed_times_keys=keys_t))
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
            return {'value': reduce_sum(float(got)) / dt_max, 'unit': UNIT, 'batches': nbl,
                    'us_per_batch': dt_max / nbl * 1e6}

        loader_api = loader_run(False)
        loader_api['api'] = ('DGDataLoader(batch_size=200) + HookManager + default-constructed '
                             'RecencyNeighborHook: outputs stay on the device, one DGBatch per '
                             'iteration (Python-bound)')
        loader_api['with_negatives'] = dict(
            loader_run(True), api='the same with RandomNegativeEdgeSamplerHook in front and seeds '
                                  '[edge_src | edge_dst | neg] (negatives drawn one window ahead)')

    # ---- the one exchange of the path: TGN node-memory join at shard boundaries (N > 1) ---------
    collective = None
    if world > 1 and not a.no_collective:
        from tgm_b200.parallel import bench_memory_join
        collective = bench_memory_join(N, 100, src[full.batch_lo * bs:min(full.batch_hi * bs, E)],
                                       dst[full.batch_lo * bs:min(full.batch_hi * bs, E)],
                                       a.join_edges, dev, reps=5)

    # ---- the reference's own eager device='cuda' mode, restated (rank 0, N=1 only) ---------------
    eager = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        eager = time_eager_cuda(a, csr, src, dst, t, x, dev)

    # ---- CPU baseline on this host (rank 0, N=1 only) ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        # steady state like the GPU arm: the `warm` edges preceding the sample are pushed first
        warm = steady_start_edge(a)
        n_cpu = min(a.cpu_sample_edges, E - warm)
        stream = tuple(None if v is None else v[:warm + n_cpu].cpu().numpy() for v in (src, dst, t, x))
        v, s, dt, warm_s = cpu_port_steady(a, stream, warm, n_cpu, 1, 0)
        cpu = {'value': v, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'host_cores': os.cpu_count(),
               'sample': f'edges [{warm}, {warm + n_cpu}) of the same stream ({s} sampled edges, '
                         f'{dt:.1f} s) after pushing the first {warm} ({warm_s:.0f} s untimed: '
                         f'steady state, full rings); C port of the reference ring sampler '
                         f'(oracle/recency_ring.c), batch by batch, 1 thread'}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': total_ms_max / a.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'int32/int64 ids+times, f32 feature copy', 'data': 'synthetic',
            'config': workload_config(a, world), 'gpu_launches': a.steps,
            'stream_edges_per_s': value / (2 * k),
            'roofline': roofline, 'cpu_baseline': cpu, 'eager_cuda_baseline': eager, 'e2e': e2e,
            'full_pass': full_pass,
            'loader_api': loader_api, 'collective': collective,
            'clocks': clocks.result(), 'numa': numa,
            'build_s': build_max,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + '\n').encode())
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
