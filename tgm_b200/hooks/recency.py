"""RecencyNeighborHook: drop-in for tgm/hooks/neighbors/recency.py:18-416 on the B200.

Same constructor, ValueErrors, requires/produces, batch attributes, dtypes, padding and
query-before-update order; the per-node ring buffers live in HBM behind a `tgm_recency` handle
(include/tgm_b200.h) and each hop is ONE kernel launch (`tgm_recency_query`) instead of ~25
eager ops plus an O(N*B) `.min()` scan + host sync (recency.py:242); the push is two launches
(`tgm_recency_update`) instead of an argsort and ~15 eager ops (recency.py:323-399).

Deliberate difference: the reference's update sort key overflows int32 when
num_nodes * (t_max + 1) >= 2**31 (recency.py:347-348) and its output is then undefined; this
implementation always follows the ideal semantics (what the reference computes inside the
parity domain, and what it would compute with `node_ids.long()`).
"""
from __future__ import annotations

import ctypes
import warnings
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from tgm_b200 import _cabi
from tgm_b200.constants import PADDED_NODE_ID
from tgm_b200.core.storage import block_views
from tgm_b200.hooks.base import SeedableHook, StatefulHook
from tgm_b200.hooks.hook_manager import register_hook_class

# seed attributes that are views of the validated store slabs (ids in [0, N), times >= 0 by
# DGData construction): checking them again would cost a host sync per batch
_TRUSTED_NODE_KEYS = ('edge_src', 'edge_dst')
_TRUSTED_TIME_KEYS = ('edge_time',)
_DEFAULT_WINDOW_BATCHES = 1024


def _is_store_view(store, tensor: Tensor) -> bool:
    """True when `tensor` aliases one of the store's validated device slabs."""
    if not tensor.is_cuda:
        return False
    base = tensor.untyped_storage().data_ptr()
    for name in ('_src', '_dst', '_t'):
        slab = getattr(store, name, None)
        if slab is not None and slab.untyped_storage().data_ptr() == base:
            return True
    return False


def _slab_offset(slab: Optional[Tensor], tensor, n: int) -> Optional[int]:
    """Row offset of `tensor` inside `slab` when it is a contiguous n-row view of that very slab
    (same storage, dtype and row shape); None otherwise."""
    if slab is None or not isinstance(tensor, Tensor) or not tensor.is_cuda or \
            tensor.dtype != slab.dtype or tensor.shape[0] != n or \
            tensor.shape[1:] != slab.shape[1:] or not tensor.is_contiguous() or \
            tensor.untyped_storage().data_ptr() != slab.untyped_storage().data_ptr():
        return None
    row_bytes = slab.element_size() * max(1, slab[0].numel()) if slab.numel() else slab.element_size()
    off = tensor.data_ptr() - slab.data_ptr()
    if off < 0 or off % row_bytes:
        return None
    return off // row_bytes


@register_hook_class
class RecencyNeighborHook(StatefulHook, SeedableHook):
    """Load the most recent neighbors of each seed node; every node keeps a fixed number of
    recent neighbors (recency sampling, k-hop, historical)."""

    _cls_requires = {'edge_src', 'edge_dst', 'edge_time'}
    _cls_produces = {'seed_nids', 'seed_times', 'nbr_nids', 'nbr_edge_time', 'nbr_edge_x',
                     'seed_node_nbr_mask'}

    def __init__(self, num_nodes: int, num_nbrs: List[int], seed_nodes_keys: List[str],
                 seed_times_keys: List[str], directed: bool = False,
                 id: Optional[str] = None, window_batches: Optional[int] = None,
                 lazy_edge_x: bool = False) -> None:
        """`window_batches` (an addition to the reference signature) controls pre-sampling: while
        the loader walks one device store front to back in equal event batches, the neighbourhoods
        of `window_batches` upcoming batches are sampled by ONE launch per hop over the stateless
        adjacency (tgm_csr_*) and each call only slices views.  Outputs are identical; any other
        call pattern (another store, a skipped batch, time-unit batches) hands the state over to
        the ring kernels.  None (default) = 1024 batches when the adjacency fits comfortably in
        free HBM; 0 = always drive the ring kernels batch by batch.
        `lazy_edge_x=True` (windowed mode only): `batch.nbr_edge_x[h]` are `LazyEdgeRows` -- edge
        ids into the store's feature table, which the fused attention reads in place -- instead of
        (S, k, D) blocks; they behave as tensors and materialise on first ordinary use."""
        if not len(num_nbrs):
            raise ValueError('num_nbrs must be non-empty')
        if not all(isinstance(x, int) and x > 0 for x in num_nbrs):
            raise ValueError('Each value in num_nbrs must be a positive integer')
        if len(seed_nodes_keys) != len(seed_times_keys):
            raise ValueError(
                f'len(seed_nodes_keys) ({len(seed_nodes_keys)}) != len(seed_times_keys) '
                f'({len(seed_times_keys)})\nseed_nodes_keys={seed_nodes_keys}, '
                f'seed_times_keys={seed_times_keys}')
        self._num_nodes = num_nodes
        self._num_nbrs = num_nbrs
        self._max_nbrs = max(num_nbrs)
        self._directed = directed
        self._seed_nodes_keys = seed_nodes_keys
        self._seed_times_keys = seed_times_keys
        self._warned_seed_None = False
        if window_batches is not None and (not isinstance(window_batches, int) or
                                           window_batches < 0):
            raise ValueError('window_batches must be a non-negative integer or None')
        self._lazy_edge_x = bool(lazy_edge_x)
        self._window_auto = window_batches is None
        self._window_batches = _DEFAULT_WINDOW_BATCHES if window_batches is None else window_batches
        self._win = None  # windowed-mode cursor, see _windowed_call
        self._standard_seeds = (list(seed_nodes_keys) == ['edge_src', 'edge_dst'] and
                                list(seed_times_keys) == ['edge_time', 'edge_time'])
        keys = list(zip(seed_nodes_keys, seed_times_keys))
        self._extra_keys = (keys[2:] if keys[:2] == [('edge_src', 'edge_time'),
                                                     ('edge_dst', 'edge_time')] else None)
        self._num_nbrs_c = (ctypes.c_int32 * len(num_nbrs))(*num_nbrs)
        self._handle = ctypes.c_void_p()   # tgm_recency*, created on first call
        self._device: Optional[torch.device] = None
        self._edge_x_dim: Optional[int] = None
        self._init_hook(id=id, seed_keys=seed_nodes_keys)

    def __del__(self, _destroy=_cabi.lib.tgm_recency_destroy) -> None:
        h = getattr(self, '_handle', None)
        if h is not None and h.value:
            _destroy(h)
            h.value = None

    @property
    def num_nbrs(self) -> List[int]:
        return self._num_nbrs

    def reset_state(self) -> None:
        self._win = None  # the next call may start a fresh windowed run
        if self._handle.value:
            _cabi.check(_cabi.lib.tgm_recency_reset(self._handle, _cabi.current_stream(self._device)))

    # -- state ------------------------------------------------------------------------------
    def _ensure_state(self, dg, ring: bool = True) -> None:
        """First call fixes D = dg.edge_x_dim or 0 and the device (recency.py:401-416); the ring
        buffers are allocated when the ring kernels are first needed (a windowed run never touches
        them until it hands over)."""
        if self._handle.value:
            return
        if self._device is not None:
            if ring:
                _cabi.check(_cabi.lib.tgm_recency_create(
                    ctypes.byref(self._handle), self._num_nodes, self._max_nbrs,
                    self._edge_x_dim, self._device.index))
            return
        device = dg.device
        if device.type != 'cuda':
            raise _cabi.TGMNativeError(
                -3, f'RecencyNeighborHook needs a CUDA DGraph, got device={device} '
                    '(tgm_b200 has no CPU fallback)')
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self._device = device
        self._edge_x_dim = dg.edge_x_dim or 0
        if ring:
            _cabi.check(_cabi.lib.tgm_recency_create(ctypes.byref(self._handle), self._num_nodes,
                                                     self._max_nbrs, self._edge_x_dim, device.index))

    def state_tensors(self) -> Dict[str, Tensor]:
        """Copies of the ring state (ids, times, feats, write_pos) for inspection/checkpoints.
        During a windowed run the state every edge served so far would have left is exported."""
        if self._device is None:
            raise RuntimeError('hook state is created on the first call')
        self._ensure_state(None)
        if isinstance(self._win, dict):
            _cabi.check(_cabi.lib.tgm_csr_export_ring(
                self._win['csr'].handle, self._win['next'], self._handle,
                _cabi.current_stream(self._device)))
        N, B, D = self._num_nodes, self._max_nbrs, self._edge_x_dim
        p = [ctypes.c_void_p() for _ in range(4)]
        _cabi.check(_cabi.lib.tgm_recency_state(self._handle, *[ctypes.byref(q) for q in p]))
        shapes = [((N, B), torch.int32), ((N, B), torch.int64), ((N, B, D), torch.float32),
                  ((N,), torch.int32)]
        out = {}
        for name, q, (shape, dtype) in zip(('ids', 'times', 'feats', 'write_pos'), p, shapes):
            if q.value and all(shape):
                out[name] = _cabi.device_view(q.value, shape, dtype, self._device).clone()
            else:
                out[name] = torch.zeros(shape, dtype=dtype, device=self._device)
        return out

    # -- the hook ---------------------------------------------------------------------------
    # -- windowed (pre-sampled) mode ----------------------------------------------------------
    def _batch_range(self, dg, batch):
        """[lo, hi) of the batch in the store's edge index space, or None if it is not a plain
        contiguous slab of a device store."""
        store = getattr(dg, '_storage', None)
        if store is None or not hasattr(store, 'edge_range') or getattr(store, 'device', None) is None:
            return None
        src = batch.edge_src
        slab = getattr(batch, '_slab', None)
        if slab is not None and slab[0] is store and src is slab[3] and \
                batch.edge_dst is slab[4] and batch.edge_time is slab[5]:
            return store, slab[1], slab[2]  # the loader's own views, untouched since
        if not isinstance(src, Tensor) or src.ndim != 1:
            return None
        n = src.numel()
        lo = _slab_offset(store._src, src, n)
        # every edge tensor must be the same rows of its own slab (a hook earlier in the chain
        # may have replaced or cast one of them)
        if lo is None or _slab_offset(store._dst, batch.edge_dst, n) != lo or \
                _slab_offset(store._t, batch.edge_time, n) != lo:
            return None
        return store, lo, lo + n

    def _published_window(self, store, batch, lo: int, n: int, node_attr: str, time_attr: str):
        """The seed window a producer hook published for `node_attr` (negatives drawn ahead for a
        stretch of the stream, hooks/negatives.py), provided the batch's attributes still are its
        views for this batch; else None."""
        pubs = getattr(batch, '_seed_windows', None)
        sw = pubs.get(node_attr) if pubs else None
        if sw is None or sw.store is not store or not (sw.e_lo <= lo and lo + n <= sw.e_hi):
            return None
        sn, stt = getattr(batch, node_attr, None), getattr(batch, time_attr, None)
        if sw.batch_size:  # the producer kept the views it handed out: identity says untouched
            j, r = divmod(lo - sw.e_lo, sw.batch_size)
            if r == 0 and sn is sw.node_views[j] and stt is sw.time_views[j]:
                return sw
        if not (isinstance(sn, Tensor) and isinstance(stt, Tensor)) or sn.numel() != n or \
                stt.numel() != n or sn.data_ptr() != sw.nodes.data_ptr() + 4 * (lo - sw.e_lo) or \
                stt.data_ptr() != sw.times.data_ptr() + 8 * (lo - sw.e_lo):
            return None
        return sw

    def _window_feasible(self, store) -> bool:
        """Default (auto) mode only: the adjacency of the whole store (entries, anchors, colocated
        feature rows, sort workspace) must fit in a quarter of the free HBM."""
        known = getattr(self, '_feasible_store', None)
        if known is not None and known[0] is store:  # asked once per store, not once per epoch
            return known[1]
        E, D = store.num_edges, self._edge_x_dim
        need = 2 * E * (16 + 8 + 16 + 4 * D)
        ok = need <= self._free_bytes() // 4
        self._feasible_store = (store, ok)
        return ok

    def _free_bytes(self) -> int:
        """Free HBM as this process sees it: what the driver reports plus what torch's allocator
        holds without using.  cudaMemGetInfo costs milliseconds (tens, now and then), so the figure
        is kept for 30 s: an epoch of a small graph is shorter than one such call."""
        import time
        now = time.monotonic()
        cached = getattr(self, '_free_bytes_cache', None)
        if cached is not None and now - cached[0] < 30.0:
            return cached[1]
        free, _ = torch.cuda.mem_get_info(self._device)
        free += torch.cuda.memory_reserved(self._device) - torch.cuda.memory_allocated(self._device)
        self._free_bytes_cache = (now, int(free))
        return int(free)

    def _windowed_call(self, dg, batch):
        """Returns the decorated batch, or None when this call cannot be served from a window (the
        caller then continues on the ring kernels, after `_leave_window` moved the state over)."""
        extra = self._extra_keys  # seed keys beyond [edge_src | edge_dst] (None: not windowable)
        if extra is None:
            return None
        slab = getattr(batch, '_slab', None)
        if slab is not None and slab[0] is getattr(dg, '_storage', None) and \
                batch.edge_src is slab[3] and batch.edge_dst is slab[4] and \
                batch.edge_time is slab[5]:
            store, lo, hi = slab[0], slab[1], slab[2]  # the loader's own views, untouched since
        else:
            rng = self._batch_range(dg, batch)
            if rng is None or rng[2] == rng[1]:
                return None
            store, lo, hi = rng
        n = hi - lo
        pub = None
        if len(extra) == 1:
            # inlined fast case of _published_window: the producer's own views, by identity
            pubs = getattr(batch, '_seed_windows', None)
            sw = pubs.get(extra[0][0]) if pubs else None
            if sw is not None and sw.store is store and sw.batch_size and sw.e_lo <= lo and \
                    hi <= sw.e_hi:
                jj, r = divmod(lo - sw.e_lo, sw.batch_size)
                if r == 0 and getattr(batch, extra[0][0], None) is sw.node_views[jj] and \
                        getattr(batch, extra[0][1], None) is sw.time_views[jj]:
                    pub = sw
            if pub is None and sw is not None:
                pub = self._published_window(store, batch, lo, n, *extra[0])
        w = self._win
        if w is not None and w.get('mode') == 'time':
            return None  # a time-window run: this batch does not continue it
        if w is None:
            if self._win is False:  # already handed over to the ring since the last reset
                return None
            if self._window_auto and not self._window_feasible(store):
                return None
            from tgm_b200.sampler import RecencyCSR
            bs = hi - lo
            cache = store._node_cache
            key = ('recency_csr', lo, bs, self._directed)
            if key not in cache:
                cache[key] = RecencyCSR(store, bs, directed=self._directed, colocate_x=True,
                                        e_start=lo)
            w = self._win = {'store': store, 'csr': cache[key], 'bs': bs, 'next': lo,
                             'w_lo': lo, 'w_hi': lo, 'hops': None, 'start': lo, 'pub': None,
                             'mask': None}
        elif w['store'] is not store or lo != w['next'] or \
                (hi - lo != w['bs'] and hi != store.num_edges):
            return None
        csr, bs = w['csr'], w['bs']
        if hi > w['w_hi']:  # pre-sample the next window, one launch per hop
            P = 3 if pub is not None else 2
            # cudaMemGetInfo costs milliseconds: the budget is fixed when a run starts
            if w.get('budget', (0, 0))[0] != P:
                w['budget'] = (P, self._window_budget_batches(bs, P))
            nwin = min(self._window_batches, w['budget'][1])
            w['w_lo'], w['w_hi'] = lo, min(lo + nwin * bs, store.num_edges)
            neg = None
            if pub is not None:
                w['w_hi'] = min(w['w_hi'], pub.e_hi)
                neg = pub.nodes[lo - pub.e_lo:w['w_hi'] - pub.e_lo]
                if not (0 <= pub.low and pub.high <= self._num_nodes):
                    self._validate([(extra[0][0], neg, True)])  # one sync per window
            w['pub'] = pub
            w['hops'] = csr.sample_window(w['w_lo'], w['w_hi'], self._num_nbrs, neg=neg,
                                          lazy_edge_x=self._lazy_edge_x)
            # per-batch views of the whole window in five C++ calls per hop (Tensor.split) rather
            # than five slicing calls per hop per batch; only the stream's last batch can be short
            rows, split = P * bs, [[], [], [], [], []]  # split[attribute][hop] = per-batch views
            for hop in w['hops']:
                for i, v in enumerate((hop.seed_nids, hop.seed_times, hop.nbr_nids,
                                       hop.nbr_edge_time, hop.nbr_edge_x)):
                    split[i].append(block_views(v, rows) if type(v) is Tensor else v.split(rows))
                rows *= hop.nbr_nids.shape[1]
            w['split'] = split
            w['mask'] = None
        elif w['pub'] is not pub:
            # the window was sampled with (without) a published seed window this batch does not
            # (does) carry any more: its rows do not describe this batch
            return None
        # rows of this batch inside the window block, hop by hop
        j = (lo - w['w_lo']) // bs
        split = w['split']
        dev = self._device
        mask = w['mask']
        if mask is None or mask[0] != n:
            mask = w['mask'] = (n, self._arange(0, n), self._arange(n, 2 * n),
                                self._arange(2 * n, 3 * n))
        if pub is not None:
            mask = {'edge_src': mask[1], 'edge_dst': mask[2], extra[0][0]: mask[3]}
            extra = []
        else:
            mask = {'edge_src': mask[1], 'edge_dst': mask[2]}
        if not extra:
            w['next'] = hi
            if self._id is None:  # the common case, without the per-attribute call overhead
                batch.seed_nids = [v[j] for v in split[0]]
                batch.seed_times = [v[j] for v in split[1]]
                batch.nbr_nids = [v[j] for v in split[2]]
                batch.nbr_edge_time = [v[j] for v in split[3]]
                batch.nbr_edge_x = [v[j] for v in split[4]]
                batch.seed_node_nbr_mask = mask
                return batch
        parts = [tuple(split[i][h][j] for i in range(5)) for h in range(len(self._num_nbrs))]
        if extra:  # seeds the window cannot know in advance: one launch per hop
            xs, xt, offset = [], [], 2 * n
            to_check = []
            for node_attr, time_attr in extra:
                for name in (node_attr, time_attr):
                    if not hasattr(batch, name):
                        raise ValueError(f'Missing seed attributes {[name]} on batch')
                sn, stt = getattr(batch, node_attr), getattr(batch, time_attr)
                if sn is None or stt is None:
                    if not self._warned_seed_None:
                        warnings.warn(
                            f'Seed attribute {node_attr if sn is None else time_attr} is None on '
                            'this batch, skipping this batch. Future occurrences will also be '
                            'skipped but the warning will be suppressed', UserWarning)
                        self._warned_seed_None = True
                    continue
                for name, tensor in ((node_attr, sn), (time_attr, stt)):
                    if not isinstance(tensor, Tensor):
                        raise ValueError(f'{name} must be a Tensor, got {type(tensor)}')
                    if tensor.ndim != 1:
                        raise ValueError(f'{name} must be 1-D, got shape {tensor.shape}')
                to_check += [(node_attr, sn, True), (time_attr, stt, False)]
                xs.append(sn.to(device=dev, dtype=torch.int32))
                xt.append(stt.to(device=dev, dtype=torch.int64))
                mask[node_attr] = self._arange(offset, offset + sn.shape[0])
                offset += sn.shape[0]
            self._validate(to_check)
            if xs:
                seeds, times = torch.cat(xs).contiguous(), torch.cat(xt).contiguous()
                cut = torch.full((1,), lo, dtype=torch.int64, device=dev)
                group = seeds.numel()
                merged = []
                for h, k in enumerate(self._num_nbrs):
                    nid, nt, nx = csr.sample(seeds, times, cut, k, self._max_nbrs, cut_group=group)
                    p = parts[h]
                    merged.append((torch.cat([p[0], seeds]), torch.cat([p[1], times]),
                                   torch.cat([p[2], nid]), torch.cat([p[3], nt]),
                                   torch.cat([p[4], nx])))
                    seeds, times = nid.reshape(-1), nt.reshape(-1)
                    group *= k
                parts = merged
        w['next'] = hi
        for i, name in enumerate(('seed_nids', 'seed_times', 'nbr_nids', 'nbr_edge_time',
                                  'nbr_edge_x')):
            self.add_batch_attribute(batch, name, [p[i] for p in parts])
        self.add_batch_attribute(batch, 'seed_node_nbr_mask', mask)
        return batch

    # -- windowed mode for time-window batches ------------------------------------------------
    def _emit_empty(self, dg, batch):
        """No seeds on this batch: empty CPU tensors with exact dtypes per hop (recency.py:127-137)."""
        hops = len(self._num_nbrs)
        add = self.add_batch_attribute
        add(batch, 'seed_nids', [torch.empty(0, dtype=torch.int32) for _ in range(hops)])
        add(batch, 'seed_times', [torch.empty(0, dtype=torch.int64) for _ in range(hops)])
        add(batch, 'nbr_nids', [torch.empty(0, dtype=torch.int32) for _ in range(hops)])
        add(batch, 'nbr_edge_time', [torch.empty(0, dtype=torch.int64) for _ in range(hops)])
        add(batch, 'nbr_edge_x', [torch.empty(0, dg.edge_x_dim).float() for _ in range(hops)])
        mask = {}
        for name in self._seed_nodes_keys:
            v = getattr(batch, name, None)
            if isinstance(v, Tensor):
                mask[name] = torch.arange(0, 0, device=v.device)
        add(batch, 'seed_node_nbr_mask', mask)
        return batch

    def _windowed_time_call(self, dg, batch, plan, j: int):
        """Time-window batches (tgm/data/loader.py:101-156) served from pre-sampled windows.  The
        loader published its plan (edge bounds of every batch); since batches are disjoint time
        ranges, a node's push order (batch, time, side, edge) is simply (time, side, edge): the
        adjacency built with ONE batch spanning the stream, with a history cut per seed = the
        first edge of its batch.  Seeds [src | dst] only; anything else returns None (ring path)."""
        if self._extra_keys != [] or plan.store is not getattr(dg, '_storage', None):
            return None
        store, bounds = plan.store, plan.bounds
        lo, hi = bounds[j], bounds[j + 1]
        n = hi - lo
        slab = getattr(batch, '_slab', None)
        if n and not (slab is not None and slab[0] is store and batch.edge_src is slab[3] and
                      batch.edge_dst is slab[4] and batch.edge_time is slab[5]):
            return None
        w = self._win
        if w is None:
            if self._window_auto and not self._window_feasible(store):
                return None
            from tgm_b200.sampler import RecencyCSR
            cache = store._node_cache
            key = ('recency_csr_time', plan.e_start, self._directed)
            if key not in cache:
                cache[key] = RecencyCSR(store, max(1, store.num_edges), directed=self._directed,
                                        colocate_x=True, e_start=plan.e_start)
            w = self._win = {'mode': 'time', 'store': store, 'plan': plan, 'csr': cache[key],
                             'next_j': j, 'j_lo': j, 'j_hi': j, 'split': None, 'next': lo}
        elif w.get('mode') != 'time' or w['plan'] is not plan or j != w['next_j']:
            return None
        if j >= w['j_hi']:  # pre-sample the next window of batches
            dev = self._device
            nb = len(bounds) - 1
            j_hi = min(j + self._window_batches, nb)
            per_edge = 0
            rows = 2
            for k in self._num_nbrs:
                per_edge += rows * k * (16 if self._lazy_edge_x else 12 + 4 * self._edge_x_dim)
                rows *= k
            if 'budget' not in w:
                w['budget'] = min(self._free_bytes() // 4, 16 << 30)
            while j_hi > j + 1 and (bounds[j_hi] - lo) * per_edge > w['budget']:
                j_hi = j + max(1, (j_hi - j) // 2)
            e_lo, e_hi = lo, bounds[j_hi]
            sizes = [bounds[i + 1] - bounds[i] for i in range(j, j_hi)]
            split = None
            if e_hi > e_lo:
                bd = plan.bounds_dev[j:j_hi + 1]
                sz = bd[1:] - bd[:-1]
                batch_of = torch.repeat_interleave(torch.arange(j_hi - j, device=dev), sz)
                start_of = bd[batch_of]                                  # first edge of the batch
                e_idx = torch.arange(e_lo, e_hi, device=dev)
                row_src = (e_idx - e_lo) + (start_of - e_lo)             # 2*(start-e_lo) + offset
                row_dst = row_src + sz[batch_of]
                S0 = 2 * (e_hi - e_lo)
                seeds = torch.empty(S0, dtype=torch.int32, device=dev)
                times = torch.empty(S0, dtype=torch.int64, device=dev)
                cut = torch.empty(S0, dtype=torch.int64, device=dev)
                es, ed, et = store._src[e_lo:e_hi], store._dst[e_lo:e_hi], store._t[e_lo:e_hi]
                seeds[row_src], seeds[row_dst] = es, ed
                times[row_src], times[row_dst] = et, et
                cut[row_src], cut[row_dst] = start_of, start_of
                csr, B = w['csr'], self._max_nbrs
                lazy = self._lazy_edge_x and self._edge_x_dim > 0 and B <= 32
                split, group, rows = [[], [], [], [], []], 1, [2 * v for v in sizes]
                for k in self._num_nbrs:
                    if lazy:
                        from tgm_b200.sampler import LazyEdgeRows
                        nid, nt, eid = csr.sample_ids(seeds, times, cut, k, B, cut_group=group)
                        nx = LazyEdgeRows(store._x, eid)
                    else:
                        nid, nt, nx = csr.sample(seeds, times, cut, k, B, cut_group=group)
                    for i, v in enumerate((seeds, times, nid, nt, nx)):
                        split[i].append(v.split(rows))  # uneven time windows: a list of sizes
                    seeds, times = nid.reshape(-1), nt.reshape(-1)
                    group *= k
                    rows = [r * k for r in rows]
            w['j_lo'], w['j_hi'], w['split'] = j, j_hi, split
        w['next_j'] = j + 1
        w['next'] = hi
        if n == 0:
            return self._emit_empty(dg, batch)
        i = j - w['j_lo']
        split = w['split']
        add = self.add_batch_attribute
        for a, name in enumerate(('seed_nids', 'seed_times', 'nbr_nids', 'nbr_edge_time',
                                  'nbr_edge_x')):
            add(batch, name, [v[i] for v in split[a]])
        add(batch, 'seed_node_nbr_mask', {'edge_src': self._arange(0, n),
                                          'edge_dst': self._arange(n, 2 * n)})
        return batch

    def _window_budget_batches(self, bs: int, seeds_per_edge: int = 2) -> int:
        """How many batches fit the pre-sampling budget (a quarter of the free HBM, at most 16 GB):
        multi-hop outputs grow as prod(k) -- 165 MB per batch for k=[20,20] at D=172."""
        per_batch, seeds = 0, seeds_per_edge * bs
        row = 16 if self._lazy_edge_x else 12 + 4 * self._edge_x_dim
        for k in self._num_nbrs:
            per_batch += seeds * k * row
            seeds *= k
        budget = min(self._free_bytes() // 4, 16 << 30)
        return max(1, int(budget // max(per_batch, 1)))

    def _arange(self, a: int, b: int) -> Tensor:
        """View [a, b) of a cached device arange (seed_node_nbr_mask rows, recency.py:221-224);
        read-only by convention, like every other view the hook hands out."""
        cache = getattr(self, '_arange_cache', None)
        if cache is None or cache.numel() < b:
            cache = torch.arange(0, max(b, 4096), device=self._device)
            self._arange_cache = cache
        return cache[a:b]

    def _leave_window(self) -> None:
        """Hand the state of a windowed run over to the ring: after this the ring holds what the
        batch-by-batch hook would hold once every edge before `next` was pushed."""
        w = self._win
        self._ensure_state(None)  # the ring buffers, if this run never needed them so far
        if isinstance(w, dict):
            _cabi.check(_cabi.lib.tgm_csr_export_ring(w['csr'].handle, w['next'], self._handle,
                                                      _cabi.current_stream(self._device)))
        self._win = False

    def __call__(self, dg, batch):
        if self._device is None:
            self._ensure_state(dg, ring=False)
        if self._window_batches and self._win is not False:
            plan = getattr(batch, '_plan', None)
            out = self._windowed_call(dg, batch) if plan is None else \
                self._windowed_time_call(dg, batch, plan[0], plan[1])
            if out is not None:
                return out
            self._leave_window()
        self._ensure_state(dg)
        if self._standard_seeds and self._step_call(dg, batch):
            return batch
        seed_nodes, seed_times, seed_mask = self._get_seed_tensors(dg, batch)
        seeds_out: List[Tensor] = []
        times_out: List[Tensor] = []
        nids: List[Tensor] = []
        nts: List[Tensor] = []
        nxs: List[Tensor] = []
        if not seed_nodes.numel():
            for _ in self._num_nbrs:  # recency.py:127-137: empty CPU tensors, exact dtypes
                seeds_out.append(torch.empty(0, dtype=torch.int32))
                times_out.append(torch.empty(0, dtype=torch.int64))
                nids.append(torch.empty(0, dtype=torch.int32))
                nts.append(torch.empty(0, dtype=torch.int64))
                nxs.append(torch.empty(0, dg.edge_x_dim).float())
        else:
            stream = _cabi.current_stream(self._device)
            for hop, k in enumerate(self._num_nbrs):
                if hop > 0:
                    seed_nodes = nids[hop - 1].flatten()
                    seed_times = nts[hop - 1].flatten()
                nid, nt, nx = self._query(seed_nodes, seed_times, k, stream)
                seeds_out.append(seed_nodes)
                times_out.append(seed_times)
                nids.append(nid)
                nts.append(nt)
                nxs.append(nx)
            if batch.edge_src.numel():
                self._update(batch, stream)
        self.add_batch_attribute(batch, 'seed_nids', seeds_out)
        self.add_batch_attribute(batch, 'seed_times', times_out)
        self.add_batch_attribute(batch, 'nbr_nids', nids)
        self.add_batch_attribute(batch, 'nbr_edge_time', nts)
        self.add_batch_attribute(batch, 'nbr_edge_x', nxs)
        self.add_batch_attribute(batch, 'seed_node_nbr_mask', seed_mask)
        return batch

    def _step_call(self, dg, batch) -> bool:
        """The whole hook call in ONE library call (tgm_recency_step) for the standard seed keys
        [edge_src, edge_dst] x [edge_time, edge_time] when the batch tensors are views of the
        validated device store (no per-key validation, concatenation or dtype moves needed).
        Returns False when the batch does not qualify; the general path then handles it."""
        src = batch.edge_src
        n = src.numel()
        store = getattr(dg, '_storage', None)
        if n == 0 or getattr(store, 'num_nodes_global', 1 << 62) > self._num_nodes or \
                self._batch_range(dg, batch) is None:
            return False
        x = batch.edge_x
        D, dev = self._edge_x_dim, self._device
        if D and not (isinstance(x, Tensor) and x.is_cuda and x.dtype == torch.float32 and
                      x.is_contiguous() and x.shape == (n, D)):
            return False
        hops = len(self._num_nbrs)
        seeds0 = torch.empty(2 * n, dtype=torch.int32, device=dev)
        times0 = torch.empty(2 * n, dtype=torch.int64, device=dev)
        nids, nts, nxs, S = [], [], [], 2 * n
        for k in self._num_nbrs:
            nids.append(torch.empty((S, k), dtype=torch.int32, device=dev))
            nts.append(torch.empty((S, k), dtype=torch.int64, device=dev))
            nxs.append(torch.empty((S, k, D), dtype=torch.float32, device=dev))
            S *= k
        P = ctypes.c_void_p * hops
        _cabi.check(_cabi.lib.tgm_recency_step(
            self._handle, src.data_ptr(), batch.edge_dst.data_ptr(), batch.edge_time.data_ptr(),
            x.data_ptr() if D else None, n, int(self._directed), hops, self._num_nbrs_c,
            seeds0.data_ptr(), times0.data_ptr(), P(*[v.data_ptr() for v in nids]),
            P(*[v.data_ptr() for v in nts]), P(*[v.data_ptr() if D else None for v in nxs]),
            _cabi.current_stream(dev)))
        add = self.add_batch_attribute
        add(batch, 'seed_nids', [seeds0] + [v.view(-1) for v in nids[:-1]])
        add(batch, 'seed_times', [times0] + [v.view(-1) for v in nts[:-1]])
        add(batch, 'nbr_nids', nids)
        add(batch, 'nbr_edge_time', nts)
        add(batch, 'nbr_edge_x', nxs)
        add(batch, 'seed_node_nbr_mask', {'edge_src': self._arange(0, n),
                                          'edge_dst': self._arange(n, 2 * n)})
        return True

    def _query(self, seeds: Tensor, tq: Tensor, k: int, stream: int
               ) -> Tuple[Tensor, Tensor, Tensor]:
        S, D, dev = seeds.numel(), self._edge_x_dim, self._device
        seeds = seeds.to(device=dev, dtype=torch.int32).contiguous()
        tq = tq.to(device=dev, dtype=torch.int64).contiguous()
        nid = torch.empty((S, k), dtype=torch.int32, device=dev)
        nt = torch.empty((S, k), dtype=torch.int64, device=dev)
        nx = torch.empty((S, k, D), dtype=torch.float32, device=dev)
        _cabi.check(_cabi.lib.tgm_recency_query(
            self._handle, seeds.data_ptr(), tq.data_ptr(), S, k, nid.data_ptr(), nt.data_ptr(),
            nx.data_ptr() if D else None, stream))
        return nid, nt, nx

    def _update(self, batch, stream: int) -> None:
        dev = self._device
        src = batch.edge_src.to(device=dev, dtype=torch.int32).contiguous()
        dst = batch.edge_dst.to(device=dev, dtype=torch.int32).contiguous()
        t = batch.edge_time.to(device=dev, dtype=torch.int64).contiguous()
        x = batch.edge_x
        if x is not None and self._edge_x_dim:
            x = x.to(device=dev, dtype=torch.float32).contiguous()
        else:
            x = None  # zeros are pushed (recency.py:325-329)
        _cabi.check(_cabi.lib.tgm_recency_update(
            self._handle, src.data_ptr(), dst.data_ptr(), t.data_ptr(), _cabi.ptr(x),
            src.numel(), int(self._directed), stream))

    def _get_seed_tensors(self, dg, batch) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
        """Seed assembly and validation (recency.py:173-237)."""
        device = batch.edge_src.device
        seeds: List[Tensor] = []
        times: List[Tensor] = []
        mask: Dict[str, Tensor] = {}
        to_check: List[Tuple[str, Tensor, bool]] = []
        store = getattr(dg, '_storage', None)
        n_global = getattr(store, 'num_nodes_global', None)
        ids_fit = n_global is not None and n_global <= self._num_nodes
        offset = 0
        for node_attr, time_attr in zip(self._seed_nodes_keys, self._seed_times_keys):
            missing = [a for a in (node_attr, time_attr) if not hasattr(batch, a)]
            if missing:
                raise ValueError(f'Missing seed attributes {missing} on batch')
            pair = [(node_attr, getattr(batch, node_attr)), (time_attr, getattr(batch, time_attr))]
            for name, tensor in pair:
                if tensor is None:  # e.g. a batch without this kind of event: skip the key
                    if not self._warned_seed_None:
                        warnings.warn(
                            f'Seed attribute {name} is None on this batch, skipping this batch. '
                            'Future occurrences will also be skipped but the warning will be '
                            'suppressed', UserWarning)
                        self._warned_seed_None = True
                    break
                if not isinstance(tensor, Tensor):
                    raise ValueError(f'{name} must be a Tensor, got {type(tensor)}')
                if tensor.ndim != 1:
                    raise ValueError(f'{name} must be 1-D, got shape {tensor.shape}')
                if name == node_attr:
                    if not (ids_fit and name in _TRUSTED_NODE_KEYS and
                            _is_store_view(store, tensor)):
                        to_check.append((name, tensor, True))
                    seeds.append(tensor.to(device))
                    n = tensor.shape[0]
                    mask[name] = (self._arange(offset, offset + n) if device == self._device
                                  else torch.arange(offset, offset + n, device=device))
                    offset += n
                else:
                    if not (name in _TRUSTED_TIME_KEYS and _is_store_view(store, tensor)):
                        to_check.append((name, tensor, False))
                    times.append(tensor.to(device))
        self._validate(to_check)
        if seeds and times:
            return torch.cat(seeds), torch.cat(times), mask
        return (torch.empty(0, dtype=torch.int32, device=device),
                torch.empty(0, dtype=torch.int64, device=device), mask)

    def _validate(self, items: List[Tuple[str, Tensor, bool]]) -> None:
        """Bounds checks of recency.py:208-229 with ONE host sync for all keys."""
        items = [(n, t, is_node) for n, t, is_node in items if t.numel()]
        if not items:
            return
        stats = torch.stack([torch.stack([t.min(), t.max()]).to(torch.int64) for _, t, _ in items])
        stats = stats.cpu().tolist()
        for (name, _, is_node), (lo, hi) in zip(items, stats):
            if is_node and (lo < 0 or hi >= self._num_nodes):
                raise ValueError(f'Seed nodes in {name} must satisfy 0 <= x < {self._num_nodes}, '
                                 f'got values in range [{lo}, {hi}]')
            if not is_node and lo < 0:
                raise ValueError(f'Seed times in {name} must be >= 0, got min value: {lo}')
