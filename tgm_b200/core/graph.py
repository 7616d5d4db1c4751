"""DGraph: immutable view (storage, slice, device) over the device-resident store.

Same public surface as tgm/core/graph.py:20-420 -- slice_events :110, slice_time :130 (end
exclusive, stored as end_time-1 :146-147), materialize :73-108, to :169 and the cached
properties -- but edge properties are zero-copy views of HBM slabs instead of per-access
`.to(device)` copies of CPU tensors (:232-263).
"""
from __future__ import annotations

from dataclasses import replace
from functools import cached_property
from typing import Optional, Tuple

import torch
from torch import Tensor

from tgm_b200.core import storage as _storage
from tgm_b200.core.batch import DGBatch
from tgm_b200.core.storage import DGSliceTracker
from tgm_b200.core.timedelta import TimeDeltaDG


def _opt_max(a, b):
    return b if a is None else a if b is None else max(a, b)


def _opt_min(a, b):
    return b if a is None else a if b is None else min(a, b)


class DGraph:
    """View over a temporal graph.  `device` must be a CUDA device for any edge data to be
    served; a CPU device gives a metadata-only view (counts, times, slicing) and raises on
    data access -- there is no CPU compute path."""

    def __init__(self, data, device: 'str | torch.device' = 'cpu') -> None:
        from tgm_b200.data.dg_data import DGData
        if not isinstance(data, DGData):
            raise TypeError(f'DGraph must be initialized with DGData, got {type(data)}')
        self._time_delta = data.time_delta
        self._device = torch.device(device)
        backend = _storage.DGStorage  # read through the module so set_dg_storage_backend applies
        try:
            self._storage = backend(data, device=self._device)
        except TypeError:  # a user backend with the plain DGStorageBase(data) signature
            self._storage = backend(data)
        if self._device.type == 'cuda' and getattr(self._storage, 'device', None) is not None:
            self._device = self._storage.device
        self._slice = DGSliceTracker()

    @classmethod
    def _from_storage(cls, storage, time_delta, device, slice) -> 'DGraph':
        obj = cls.__new__(cls)
        obj._storage, obj._time_delta, obj._device, obj._slice = storage, time_delta, device, slice
        return obj

    # -- views ----------------------------------------------------------------------------
    def slice_events(self, start_idx: Optional[int] = None, end_idx: Optional[int] = None
                     ) -> 'DGraph':
        if start_idx is not None and end_idx is not None and start_idx > end_idx:
            raise ValueError(f'start_idx ({start_idx}) must be <= end_idx ({end_idx})')
        s = replace(self._slice)
        s.start_idx = _opt_max(start_idx, s.start_idx)
        s.end_idx = _opt_min(end_idx, s.end_idx)
        return DGraph._from_storage(self._storage, self._time_delta, self._device, s)

    def slice_time(self, start_time: Optional[int] = None, end_time: Optional[int] = None
                   ) -> 'DGraph':
        if start_time is not None and end_time is not None and start_time > end_time:
            raise ValueError(f'start_time ({start_time}) must be <= end_time ({end_time})')
        if end_time is not None:
            end_time -= 1  # exclusive end -> inclusive bound (graph.py:146-147)
        s = replace(self._slice)
        s.start_time = _opt_max(start_time, s.start_time)
        s.end_time = _opt_min(end_time, s.end_time)
        return DGraph._from_storage(self._storage, self._time_delta, self._device, s)

    def to(self, device: 'str | torch.device') -> 'DGraph':
        """View on another device (graph.py:169-181).  Moving a host-side (metadata-only) graph to
        a CUDA device uploads the store once; the new view keeps the slice."""
        device = torch.device(device)
        storage = self._storage
        if device.type == 'cuda' and getattr(storage, 'device', 'n/a') is None and \
                getattr(storage, '_data', None) is not None:
            storage = type(storage)(storage._data, device=device)
            device = storage.device
        return DGraph._from_storage(storage, self._time_delta, device, replace(self._slice))

    def materialize(self, materialize_features: bool = True) -> DGBatch:
        src, dst, time = self._edges
        batch = DGBatch(self._mv(src), self._mv(dst), self._mv(time))
        if materialize_features:
            nx = self.node_x
            if nx is not None:
                batch.node_x_time, nids = nx._indices()
                batch.node_x_nids = nids.to(torch.int32)
                batch.node_x = nx._values()
            ex = self.edge_x
            if ex is not None:
                batch.edge_x = ex
            ny = self.node_y
            if ny is not None:
                batch.node_y_time, nids = ny._indices()
                batch.node_y_nids = nids.to(torch.int32)
                batch.node_y = ny._values()
        et = self.edge_type
        if et is not None:
            batch.edge_type = et
        return batch

    def __len__(self) -> int:
        return self.num_timestamps

    def __str__(self) -> str:
        return (f'DGraph(storage={type(self._storage).__name__}, time_delta={self.time_delta}, '
                f'device={self.device})')

    # -- metadata -------------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self._device

    @property
    def time_delta(self) -> TimeDeltaDG:
        return self._time_delta

    @cached_property
    def start_time(self) -> Optional[int]:
        if self._slice.start_time is None:
            self._slice.start_time = self._storage.get_start_time(self._slice)
        return self._slice.start_time

    @cached_property
    def end_time(self) -> Optional[int]:
        if self._slice.end_time is None:
            self._slice.end_time = self._storage.get_end_time(self._slice)
        return self._slice.end_time

    @cached_property
    def num_nodes(self) -> int:
        nodes = self._storage.get_nodes(self._slice)
        return max(nodes) + 1 if nodes else 0

    @cached_property
    def num_node_events(self) -> int:
        return len(self._node_events[1])

    @cached_property
    def num_node_labels(self) -> int:
        return len(self._node_labels[1])

    @cached_property
    def num_edge_events(self) -> int:
        lo, hi = self._storage.edge_range(self._slice) if hasattr(
            self._storage, 'edge_range') else (0, len(self._edges[2]))
        return hi - lo

    @cached_property
    def num_timestamps(self) -> int:
        return self._storage.get_num_timestamps(self._slice)

    @cached_property
    def num_events(self) -> int:
        return self._storage.get_num_events(self._slice)

    # -- data (device views) ----------------------------------------------------------------
    def _mv(self, t: Optional[Tensor]) -> Optional[Tensor]:
        return None if t is None else t if t.device == self._device else t.to(self._device)

    @cached_property
    def _edges(self) -> Tuple[Tensor, Tensor, Tensor]:
        return self._storage.get_edges(self._slice)

    @property
    def edge_src(self) -> Tensor:
        return self._mv(self._edges[0])

    @property
    def edge_dst(self) -> Tensor:
        return self._mv(self._edges[1])

    @property
    def edge_time(self) -> Tensor:
        return self._mv(self._edges[2])

    @cached_property
    def _edge_x(self) -> Optional[Tensor]:
        return self._storage.get_edge_x(self._slice)

    @property
    def edge_x(self) -> Optional[Tensor]:
        return self._mv(self._edge_x)

    @cached_property
    def _edge_type(self) -> Optional[Tensor]:
        return self._storage.get_edge_type(self._slice)

    @property
    def edge_type(self) -> Optional[Tensor]:
        return self._mv(self._edge_type)

    @cached_property
    def _node_events(self) -> Tuple[Tensor, Tensor]:
        return self._storage.get_node_events(self._slice)

    @property
    def node_x_nids(self) -> Tensor:
        return self._mv(self._node_events[0])

    @property
    def node_x_time(self) -> Tensor:
        return self._mv(self._node_events[1])

    @cached_property
    def _node_x(self) -> Optional[Tensor]:
        return self._storage.get_node_x(self._slice)

    @property
    def node_x(self) -> Optional[Tensor]:
        return self._mv(self._node_x)

    @cached_property
    def _node_labels(self) -> Tuple[Tensor, Tensor]:
        return self._storage.get_node_labels(self._slice)

    @property
    def node_y_nids(self) -> Tensor:
        return self._mv(self._node_labels[0])

    @property
    def node_y_time(self) -> Tensor:
        return self._mv(self._node_labels[1])

    @cached_property
    def _node_y(self) -> Optional[Tensor]:
        return self._storage.get_node_y(self._slice)

    @property
    def node_y(self) -> Optional[Tensor]:
        return self._mv(self._node_y)

    @cached_property
    def _static_node_x(self) -> Optional[Tensor]:
        return self._mv(self._storage.get_static_node_x())

    @property
    def static_node_x(self) -> Optional[Tensor]:
        return self._static_node_x

    @property
    def node_type(self) -> Optional[Tensor]:
        return self._mv(self._storage.get_node_type())

    @cached_property
    def static_node_x_dim(self) -> Optional[int]:
        return self._storage.get_static_node_x_dim()

    @cached_property
    def node_x_dim(self) -> Optional[int]:
        return self._storage.get_node_x_dim()

    @cached_property
    def node_y_dim(self) -> Optional[int]:
        return self._storage.get_node_y_dim()

    @cached_property
    def edge_x_dim(self) -> Optional[int]:
        return self._storage.get_edge_x_dim()
