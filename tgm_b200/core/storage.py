"""Storage engine: slice tracker, abstract storage API and the B200 device-resident backend.

API = tgm/core/_storage/base.py:10-118 (DGSliceTracker, DGStorageBase and its getters, same
names / argument meaning / return dtypes).  `DeviceCOOStorage` replaces DGStorageArrayBackend
(tgm/core/_storage/backends/array_backend.py:15-321): the time-sorted edge arrays live in HBM
behind a `tgm_store` handle (include/tgm_b200.h), a slice is two O(log E) binary searches over
the timestamps plus a pointer offset, and batch materialisation is a zero-copy
view -- the reference's per-batch O(E) boolean masks (:59,:264) and per-property H2D copies
(tgm/core/graph.py:232-263) are gone.
"""
from __future__ import annotations

import ctypes
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Optional, Set, Tuple

import numpy as np
import torch
from torch import Tensor

from tgm_b200 import _cabi
from tgm_b200.constants import PADDED_NODE_ID


def block_views(v: Tensor, n: int):
    """Views of the consecutive n-row blocks of `v` (`Tensor.split(n)` semantics: the last one may
    be shorter).  When the rows divide evenly the views come from one unflatten + unbind, which
    costs ~12 % less per view than split -- and view creation is most of the per-batch cost of the
    loader and the windowed hooks (ten views per batch at ~0.9 us each)."""
    rows = v.shape[0]
    if type(v) is Tensor and n > 0 and rows >= n and rows % n == 0:
        return v.unflatten(0, (rows // n, n)).unbind(0)
    return v.split(n)


@dataclass(slots=True)
class DGSliceTracker:
    """Time / event-index window of a view (base.py:10-17).  Times are inclusive on both ends;
    indices are [start_idx, end_idx) positions on the unified event timeline."""
    start_time: Optional[int] = None
    end_time: Optional[int] = None
    start_idx: Optional[int] = None
    end_idx: Optional[int] = None


class DGStorageBase(ABC):
    """Getter contract every storage backend honours (base.py:20-118)."""

    @abstractmethod
    def __init__(self, data) -> None: ...
    @abstractmethod
    def get_start_time(self, slice: DGSliceTracker) -> Optional[int]: ...
    @abstractmethod
    def get_end_time(self, slice: DGSliceTracker) -> Optional[int]: ...
    @abstractmethod
    def get_nodes(self, slice: DGSliceTracker) -> Set[int]: ...
    @abstractmethod
    def get_edges(self, slice: DGSliceTracker) -> Tuple[Tensor, Tensor, Tensor]: ...
    @abstractmethod
    def get_node_events(self, slice: DGSliceTracker) -> Tuple[Tensor, Tensor]: ...
    @abstractmethod
    def get_node_labels(self, slice: DGSliceTracker) -> Tuple[Tensor, Tensor]: ...
    @abstractmethod
    def get_num_timestamps(self, slice: DGSliceTracker) -> int: ...
    @abstractmethod
    def get_num_events(self, slice: DGSliceTracker) -> int: ...
    @abstractmethod
    def get_node_x(self, slice: DGSliceTracker) -> Optional[Tensor]: ...
    @abstractmethod
    def get_node_y(self, slice: DGSliceTracker) -> Optional[Tensor]: ...
    @abstractmethod
    def get_edge_x(self, slice: DGSliceTracker) -> Optional[Tensor]: ...
    @abstractmethod
    def get_edge_type(self, slice: DGSliceTracker) -> Optional[Tensor]: ...
    @abstractmethod
    def get_static_node_x(self) -> Optional[Tensor]: ...
    @abstractmethod
    def get_node_type(self) -> Optional[Tensor]: ...
    @abstractmethod
    def get_node_x_dim(self) -> Optional[int]: ...
    @abstractmethod
    def get_node_y_dim(self) -> Optional[int]: ...
    @abstractmethod
    def get_edge_x_dim(self) -> Optional[int]: ...
    @abstractmethod
    def get_static_node_x_dim(self) -> Optional[int]: ...
    @abstractmethod
    def get_nbrs(self, seed_nodes: Tensor, num_nbrs: int, slice: DGSliceTracker,
                 directed: bool, reference_rng: bool = False) -> Tuple[Tensor, ...]: ...


class DeviceCOOStorage(DGStorageBase):
    """Time-sorted COO edge store resident in HBM.

    device=None builds a metadata-only store (slice bounds, counts, times work on any host;
    touching edge data raises): there is no CPU compute fallback.
    """

    def __init__(self, data, device: 'torch.device | str | None' = None) -> None:
        self._data = data
        self._device = None if device is None else torch.device(device)
        if self._device is not None and self._device.type != 'cuda':
            self._device = None
        self._E = int(data.edge_index.shape[0])
        self._D = 0 if data.edge_x is None else int(data.edge_x.shape[1])
        self._num_nodes = int(data.num_nodes)
        self._edge_only = data.time.shape[0] == self._E  # event index == edge index
        self._time_np_cache = data.time.numpy()
        self._edge_pos_np = None if self._edge_only else data.edge_mask.numpy()
        edge_time = data.time if self._edge_only else data.time[data.edge_mask]
        self._edge_time_host = edge_time.contiguous()
        self._handle = ctypes.c_void_p()
        self._src = self._dst = self._t = self._x = self._edge_type = None
        self._node_cache: dict = {}

        src = data.edge_index[:, 0].contiguous()
        dst = data.edge_index[:, 1].contiguous()
        if self._device is None:
            _cabi.check(_cabi.lib.tgm_store_create(
                ctypes.byref(self._handle), src.data_ptr(), dst.data_ptr(),
                self._edge_time_host.data_ptr(), None if data.edge_x is None else
                data.edge_x.contiguous().data_ptr(), self._E, self._D, self._num_nodes, -1,
                _cabi.TGM_MEM_HOST, None))
            return
        _cabi.require_device()
        dev = self._device
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
            self._device = dev
        # torch owns the device slabs (plumbing); the store adopts the pointers zero-copy
        self._src = src.to(dev, non_blocking=True)
        self._dst = dst.to(dev, non_blocking=True)
        self._t = self._edge_time_host.to(dev, non_blocking=True)
        self._x = None if data.edge_x is None else data.edge_x.contiguous().to(dev, non_blocking=True)
        self._edge_type = None if data.edge_type is None else data.edge_type.to(dev)
        torch.cuda.synchronize(dev)
        _cabi.check(_cabi.lib.tgm_store_create(
            ctypes.byref(self._handle), self._src.data_ptr(), self._dst.data_ptr(),
            self._t.data_ptr(), _cabi.ptr(self._x), self._E, self._D, self._num_nodes, dev.index,
            _cabi.TGM_MEM_DEVICE, self._edge_time_host.data_ptr()))

    @classmethod
    def from_device_tensors(cls, src: Tensor, dst: Tensor, t: Tensor, x: Optional[Tensor],
                            num_nodes: int) -> 'DeviceCOOStorage':
        """Adopt an edge stream that already lives in HBM (src/dst int32[E], t int64[E] sorted,
        x float32[E,D] or None) without a host DGData mirror.  Serves the samplers and the edge
        getters; getters that need host-side node events / labels are not available."""
        if not (src.is_cuda and dst.is_cuda and t.is_cuda and (x is None or x.is_cuda)):
            raise _cabi.TGMNativeError(-1, 'from_device_tensors needs CUDA tensors')
        if src.dtype != torch.int32 or dst.dtype != torch.int32 or t.dtype != torch.int64:
            raise TypeError('src/dst must be int32 and t int64')
        if x is not None and (x.dtype != torch.float32 or x.ndim != 2 or x.shape[0] != src.numel()):
            raise TypeError('x must be float32 [E, D]')
        self = cls.__new__(cls)
        self._data = None
        self._device = src.device
        self._E, self._D = int(src.numel()), 0 if x is None else int(x.shape[1])
        self._num_nodes = int(num_nodes)
        self._edge_only = True
        # no host mirror of the timestamps (800 MB at 1e8 edges): slice bounds are searched on the
        # device; `_time_np` reads one back on first use (get_num_timestamps on a view)
        self._edge_time_host = None
        self._time_np_cache = None
        self._edge_pos_np = None
        self._node_cache = {}
        self._src, self._dst, self._t = src.contiguous(), dst.contiguous(), t.contiguous()
        self._x = None if x is None else x.contiguous()
        self._edge_type = None
        self._handle = ctypes.c_void_p()
        _cabi.check(_cabi.lib.tgm_store_create(
            ctypes.byref(self._handle), self._src.data_ptr(), self._dst.data_ptr(),
            self._t.data_ptr(), _cabi.ptr(self._x), self._E, self._D, self._num_nodes,
            self._device.index, _cabi.TGM_MEM_DEVICE, None))
        return self

    @property
    def _time_np(self):
        if self._time_np_cache is None:
            self._time_np_cache = self._t.cpu().numpy()
        return self._time_np_cache

    def __del__(self, _destroy=_cabi.lib.tgm_store_destroy) -> None:
        h = getattr(self, '_handle', None)
        if h is not None and h.value:
            _destroy(h)
            h.value = None

    # -- handles for the samplers ---------------------------------------------------------
    @property
    def handle(self) -> ctypes.c_void_p:
        return self._handle

    @property
    def device(self) -> Optional[torch.device]:
        return self._device

    @property
    def num_edges(self) -> int:
        return self._E

    @property
    def num_nodes_global(self) -> int:
        return self._num_nodes

    @property
    def _has_edge_x(self) -> bool:
        return self._x is not None if self._data is None else self._data.edge_x is not None

    def _require_device(self) -> None:
        if self._device is None:
            raise _cabi.TGMNativeError(
                -3, 'edge data lives on the GPU: construct the DGraph with device="cuda" '
                    '(tgm_b200 has no CPU fallback)')

    # -- slice bounds (array_backend.py:301-321) ------------------------------------------
    def _event_bounds(self, s: DGSliceTracker) -> Tuple[int, int]:
        """[lb, ub) on the unified event timeline."""
        idx_lo = -1 if not s.start_idx else int(s.start_idx)   # `start_idx or 0` (:319)
        idx_hi = -1 if not s.end_idx else int(s.end_idx)       # `end_idx or len(ts)` (:319-320)
        if self._edge_only:
            lb, ub = ctypes.c_int64(), ctypes.c_int64()
            _cabi.check(_cabi.lib.tgm_store_bounds(
                self._handle, int(s.start_time or 0), int(s.start_time is not None),
                int(s.end_time or 0), int(s.end_time is not None), idx_lo, idx_hi,
                ctypes.byref(lb), ctypes.byref(ub)))
            return lb.value, ub.value
        ts = self._time_np
        lo = 0 if s.start_time is None else int(np.searchsorted(ts, s.start_time, 'left'))
        hi = len(ts) if s.end_time is None else int(np.searchsorted(ts, s.end_time, 'right'))
        cl, ch = max(idx_lo, 0), (len(ts) if idx_hi < 0 else idx_hi)
        return max(cl, min(ch, lo)), max(cl, min(ch, hi))

    def edge_range(self, s: DGSliceTracker) -> Tuple[int, int]:
        """[lo, hi) in edge-index space: the slab of the slice."""
        lb, ub = self._event_bounds(s)
        if self._edge_only:
            return lb, max(lb, ub)
        pos = self._edge_pos_np
        return int(np.searchsorted(pos, lb, 'left')), int(np.searchsorted(pos, max(lb, ub), 'left'))

    def _sub_range(self, mask: Optional[Tensor], s: DGSliceTracker) -> Tuple[int, int]:
        if mask is None:
            return 0, 0
        lb, ub = self._event_bounds(s)
        m = mask.numpy()
        return int(np.searchsorted(m, lb, 'left')), int(np.searchsorted(m, max(lb, ub), 'left'))

    # -- loader fast path -------------------------------------------------------------------
    _CHUNK_BATCHES = 4096

    @property
    def edges_only(self) -> bool:
        """True when the event timeline holds edge events only and there are no edge types: a
        loader batch is then fully described by its slab [lo, hi) of the edge arrays."""
        return bool(self._edge_only and self._device is not None and
                    getattr(self, '_edge_type', None) is None and
                    (self._data is None or (self._data.node_x_mask is None and
                                            self._data.node_y_mask is None)))

    def batch_chunk(self, origin: int, batch_size: int, c: int):
        """Per-batch views of chunk `c` (4096 batches starting at origin + c * 4096 * batch_size):
        a tuple (src, dst, t, x) of `Tensor.split` tuples -- one C++ call yields the views of 4096
        batches instead of four slicing calls per batch.  One live chunk: the loader walks forward."""
        key = ('batch_views', origin, batch_size, c)
        chunk = self._node_cache.get(key)
        if chunk is None:
            for k in [k for k in self._node_cache if isinstance(k, tuple) and k[0] == 'batch_views']:
                del self._node_cache[k]
            a = origin + c * self._CHUNK_BATCHES * batch_size
            b = min(a + self._CHUNK_BATCHES * batch_size, self._E)
            chunk = tuple(None if v is None else block_views(v[a:b], batch_size)
                          for v in (self._src, self._dst, self._t, self._x))
            self._node_cache[key] = chunk
        return chunk

    def batch_views(self, origin: int, batch_size: int, lo: int, hi: int):
        """(src, dst, t, x) views of the slab [lo, hi), where lo = origin + j * batch_size.
        Identical tensors to get_edges/get_edge_x."""
        j = (lo - origin) // batch_size
        c, r = divmod(j, self._CHUNK_BATCHES)
        chunk = self.batch_chunk(origin, batch_size, c)
        if hi - lo == batch_size or lo + batch_size > self._E:
            if hi - lo == len(chunk[0][r]):
                return tuple(None if v is None else v[r] for v in chunk)
        return (self._src[lo:hi], self._dst[lo:hi], self._t[lo:hi],
                None if self._x is None else self._x[lo:hi])

    # -- getters --------------------------------------------------------------------------
    def _time_at(self, i: int) -> int:
        if self._time_np_cache is None:  # adopted device stream: read the one word
            return int(self._t[i])
        return int(self._time_np_cache[i])

    def get_start_time(self, slice: DGSliceTracker) -> Optional[int]:
        lb, ub = self._event_bounds(slice)
        return None if lb >= ub else self._time_at(lb)

    def get_end_time(self, slice: DGSliceTracker) -> Optional[int]:
        lb, ub = self._event_bounds(slice)
        return None if lb >= ub else self._time_at(ub - 1)

    def get_num_events(self, slice: DGSliceTracker) -> int:
        lb, ub = self._event_bounds(slice)
        return ub - lb

    def get_num_timestamps(self, slice: DGSliceTracker) -> int:
        lb, ub = self._event_bounds(slice)
        return int(np.unique(self._time_np[lb:ub]).size) if ub > lb else 0

    def get_nodes(self, slice: DGSliceTracker) -> Set[int]:
        lo, hi = self.edge_range(slice)
        if self._data is None:  # adopted device stream: no host mirror of the ids
            both = torch.cat([self._src[lo:hi], self._dst[lo:hi]])
            return set(torch.unique(both).cpu().tolist())
        nodes = set(np.unique(self._data.edge_index[lo:hi].numpy()).tolist())
        a, b = self._sub_range(self._data.node_x_mask, slice)
        if b > a:
            nodes.update(np.unique(self._data.node_x_nids[a:b].numpy()).tolist())
        return nodes

    def get_edges(self, slice: DGSliceTracker) -> Tuple[Tensor, Tensor, Tensor]:
        """(src int32, dst int32, time int64) DEVICE views of the slab -- no copy, no mask."""
        self._require_device()
        lo, hi = self.edge_range(slice)
        return self._src[lo:hi], self._dst[lo:hi], self._t[lo:hi]

    def get_edge_x(self, slice: DGSliceTracker) -> Optional[Tensor]:
        if not self._has_edge_x:
            return None
        self._require_device()
        lo, hi = self.edge_range(slice)
        return None if hi <= lo else self._x[lo:hi]  # None on an empty slice (:264-266)

    def get_edge_type(self, slice: DGSliceTracker) -> Optional[Tensor]:
        if self._edge_type is None and (self._data is None or self._data.edge_type is None):
            return None
        self._require_device()
        lo, hi = self.edge_range(slice)
        return None if hi <= lo else self._edge_type[lo:hi]

    def get_node_events(self, slice: DGSliceTracker) -> Tuple[Tensor, Tensor]:
        if self._data is None or self._data.node_x_mask is None:
            return torch.empty(0, dtype=torch.int32), torch.empty(0, dtype=torch.int64)
        a, b = self._sub_range(self._data.node_x_mask, slice)
        return self._data.node_x_nids[a:b], self._data.time[self._data.node_x_mask[a:b]]

    def get_node_labels(self, slice: DGSliceTracker) -> Tuple[Tensor, Tensor]:
        if self._data is None or self._data.node_y_mask is None:
            return torch.empty(0, dtype=torch.int32), torch.empty(0, dtype=torch.int64)
        a, b = self._sub_range(self._data.node_y_mask, slice)
        return self._data.node_y_nids[a:b], self._data.time[self._data.node_y_mask[a:b]]

    def _sparse_node_tensor(self, slice, mask, nids, values, dim) -> Optional[Tensor]:
        """sparse COO (time, node, dim) of the node events in the slice (:179-257)."""
        if values is None:
            return None
        a, b = self._sub_range(mask, slice)
        if b <= a:
            return None
        time = self._data.time[mask[a:b]]
        nodes = nids[a:b]
        max_node = int(nodes.max())
        lo, hi = self.edge_range(slice)
        if hi > lo:
            max_node = max(max_node, int(self._data.edge_index[lo:hi].max()))
        for om, on in ((self._data.node_x_mask, self._data.node_x_nids),
                       (self._data.node_y_mask, self._data.node_y_nids)):
            if om is not None and om is not mask:
                oa, ob = self._sub_range(om, slice)
                if ob > oa:
                    max_node = max(max_node, int(on[oa:ob].max()))
        _, ub = self._event_bounds(slice)
        max_time = slice.end_time or int(self._time_np[ub - 1])
        idx = torch.stack([time, nodes.to(torch.int64)], 0)
        return torch.sparse_coo_tensor(idx, values[a:b], (max_time + 1, max_node + 1, dim))

    def get_node_x(self, slice: DGSliceTracker) -> Optional[Tensor]:
        d = self._data
        if d is None:
            return None
        return self._sparse_node_tensor(slice, d.node_x_mask, d.node_x_nids, d.node_x,
                                        self.get_node_x_dim())

    def get_node_y(self, slice: DGSliceTracker) -> Optional[Tensor]:
        d = self._data
        if d is None:
            return None
        return self._sparse_node_tensor(slice, d.node_y_mask, d.node_y_nids, d.node_y,
                                        self.get_node_y_dim())

    def get_static_node_x(self) -> Optional[Tensor]:
        return None if self._data is None else self._data.static_node_x

    def get_node_type(self) -> Optional[Tensor]:
        return None if self._data is None else self._data.node_type

    def get_node_x_dim(self) -> Optional[int]:
        d = self._data
        return None if d is None or d.node_x is None else int(d.node_x.shape[1])

    def get_node_y_dim(self) -> Optional[int]:
        d = self._data
        return None if d is None or d.node_y is None else int(d.node_y.shape[1])

    def get_edge_x_dim(self) -> Optional[int]:
        return self._D if self._has_edge_x else None

    def get_static_node_x_dim(self) -> Optional[int]:
        sx = None if self._data is None else self._data.static_node_x
        return None if sx is None else int(sx.shape[1])

    def get_nbrs(self, seed_nodes: Tensor, num_nbrs: int, slice: DGSliceTracker,
                 directed: bool, reference_rng: bool = False) -> Tuple[Tensor, ...]:
        """Neighbours among all edges of the slice (array_backend.py:108-171), right-padded.
        Served from the (edge, side)-ordered adjacency by `tgm_csr_sample_uniform`: seeds with
        <= `num_nbrs` candidates get exactly the reference's rows; with more, a uniform subset is
        drawn on the device.  `reference_rng=True` (not part of the reference signature) instead
        replays the reference's own `random.sample` calls (:152-153) on the host from the
        candidate counts and gathers those picks on the device: bit-exact under `random.seed`,
        at the price of one host sync per call."""
        self._require_device()
        from tgm_b200.sampler import full_history_neighbors
        return full_history_neighbors(self, seed_nodes, num_nbrs, slice, directed,
                                      reference_rng=reference_rng)


# registry mirroring tgm/core/_storage/__init__.py:14-28 and backends/__init__.py:3-7
DGStorageBackends = {'DeviceCOOBackend': DeviceCOOStorage}
DGStorage = DeviceCOOStorage


def get_dg_storage_backend():
    return DGStorage


def set_dg_storage_backend(backend) -> None:
    global DGStorage
    if isinstance(backend, type) and issubclass(backend, DGStorageBase):
        DGStorage = backend
    elif isinstance(backend, str) and backend in DGStorageBackends:
        DGStorage = DGStorageBackends[backend]
    else:
        raise ValueError(f'Unrecognized DGStorage backend: {backend}, expected one of: '
                         f'{list(DGStorageBackends.keys())}')
