"""DGDataLoader: sequential batches over a DGraph, hooks applied per batch.

Same constructor arguments, batching rules and empty-batch policy as tgm/data/loader.py:64-184
(event-ordered 'r' batches are [i, i+bs) event indices :136-148,:158-160; time-unit batches are
[t, t+bs) windows; `drop_last`; on_empty in {'skip','raise',None} :20-61).  It is a plain
Python iterator rather than a torch DataLoader subclass: there are no workers to manage (state
lives on the GPU) and the torch DataLoader machinery costs more per step than the kernels do.
"""
from __future__ import annotations

from typing import Any, Iterator, Optional

from tgm_b200.core.batch import DGBatch
from tgm_b200.core.graph import DGraph
from tgm_b200.core.storage import DGSliceTracker
from tgm_b200.core.timedelta import TimeDeltaDG
from tgm_b200.exceptions import (EmptyBatchError, EventOrderedConversionError,
                                 InvalidDiscretizationError)

_ON_EMPTY = ('skip', 'raise', None)


class _TimePlan:
    """Edge-index bounds of every time-window batch of a loader over an edge-only device store."""

    _CHUNK = 4096

    def __init__(self, store, dg: DGraph, starts: range, batch_size: int) -> None:
        import torch
        lo0, hi0 = store.edge_range(dg._slice)
        th = torch.arange(starts.start, starts.start + (len(starts) + 1) * batch_size, batch_size,
                          device=store.device, dtype=torch.int64)
        b = torch.searchsorted(store._t[lo0:hi0], th) + lo0
        self.store, self.starts, self.batch_size = store, starts, batch_size
        self.bounds_dev = b                       # int64[nb + 1], non-decreasing
        self.bounds = b.tolist()                  # one host sync for the whole plan
        self.e_start = lo0
        self._chunk = None

    def views(self, j: int):
        """(src, dst, t, x) views of batch j, served from per-chunk `Tensor.split` tuples."""
        c, r = divmod(j, self._CHUNK)
        if self._chunk is None or self._chunk[0] != c:
            a = c * self._CHUNK
            bnd = self.bounds[a:a + self._CHUNK + 1]
            sizes = [y - x for x, y in zip(bnd, bnd[1:])]
            st = self.store
            self._chunk = (c, tuple(None if v is None else v[bnd[0]:bnd[-1]].split(sizes)
                                    for v in (st._src, st._dst, st._t, st._x)))
        ch = self._chunk[1]
        return ch[0][r], ch[1][r], ch[2][r], None if ch[3] is None else ch[3][r]


class DGDataLoader:
    def __init__(self, dg: DGraph, batch_size: int = 1, batch_unit: str = 'r',
                 on_empty: Optional[str] = 'skip', hook_manager=None, **kwargs: Any) -> None:
        if batch_size <= 0:
            raise ValueError(f'batch_size must be > 0 but got {batch_size}')
        if on_empty not in _ON_EMPTY:
            raise ValueError(f'Invalid on_empty={on_empty}, expected one of: {list(_ON_EMPTY)}')
        unit = TimeDeltaDG(batch_unit)
        if dg.time_delta.is_event_ordered and unit.is_time_ordered:
            raise EventOrderedConversionError(
                'Cannot iterate event-ordered dg using time-ordered batch_unit')
        if dg.time_delta.is_time_ordered and unit.is_time_ordered:
            unit = TimeDeltaDG(batch_unit, value=batch_size)
            if dg.time_delta.is_coarser_than(unit):
                raise InvalidDiscretizationError(
                    f'Tried to construct a data loader on a DGraph with time delta: '
                    f'{dg.time_delta} which is strictly coarser than the batch_unit: '
                    f'{batch_unit}, batch_size: {batch_size}.')
            batch_size = int(unit.convert(dg.time_delta))
        if dg.start_time is None or dg.end_time is None:
            raise ValueError('cannot iterate an empty DGraph')
        self._dg = dg
        self._batch_size = batch_size
        self._hook_manager = hook_manager
        self._on_empty = on_empty
        self._by_events = unit.is_event_ordered
        if self._by_events:
            self._slice_op = dg.slice_events
            start, stop = 0, dg.num_events  # loader.py:137-139
        else:
            self._slice_op = dg.slice_time
            start, stop = dg.start_time, dg.end_time + 1
        if kwargs.get('drop_last', False):
            self._starts = range(start, stop - batch_size, batch_size)
        else:
            self._starts = range(start, stop, batch_size)
        # Fast path: event-ordered batches over an edge-only device store are plain slabs
        # [max(i, lo0), min(i + bs, hi0)) of the edge arrays -- what slice_events + materialize
        # compute (graph.py:110-152, 73-108) without the per-batch view objects and bound lookups.
        self._fast = None
        store = getattr(dg, '_storage', None)
        if self._by_events and getattr(store, 'edges_only', False) and dg.device == store.device:
            self._fast = (store,) + tuple(store.edge_range(dg._slice))
            self._origin = self._fast[1] - self._fast[1] % batch_size
            self._chunk = None
        # The same for time-window batches (batch j = edges with start + j*bs <= t < start + (j+1)*bs,
        # graph.py:130-152): the window bounds of the whole plan come from ONE device searchsorted,
        # and the plan travels on every batch (`batch._plan`) so that the neighbour hook can
        # pre-sample many upcoming windows per launch.
        self._time_plan = None
        if not self._by_events and getattr(store, 'edges_only', False) and \
                dg.device == store.device and len(self._starts):
            self._time_plan = _TimePlan(store, dg, self._starts, batch_size)

    @property
    def dgraph(self) -> DGraph:
        return self._dg

    def __len__(self) -> int:
        return len(self._starts)

    def _load_time(self, start: int) -> DGBatch:
        plan = self._time_plan
        j = (start - plan.starts.start) // plan.batch_size
        lo, hi = plan.bounds[j], plan.bounds[j + 1]
        src, dst, t, x = plan.views(j)
        n = hi - lo
        batch = DGBatch(src, dst, t, x if n else None)
        if n:
            batch._slab = (plan.store, lo, hi, src, dst, t)
        batch._plan = (plan, j)
        hm = self._hook_manager
        if hm is not None:
            src_dg = self._dg
            s = src_dg._slice
            t_lo, t_hi = start, start + plan.batch_size - 1  # slice_time's inclusive bounds
            view = DGraph._from_storage(plan.store, src_dg._time_delta, src_dg._device, DGSliceTracker(
                t_lo if s.start_time is None else max(t_lo, s.start_time),
                t_hi if s.end_time is None else min(t_hi, s.end_time), s.start_idx, s.end_idx))
            batch = hm.execute_active_hooks(view, batch)
        return batch

    def _load(self, start: int) -> DGBatch:
        if self._fast is not None:
            return self._load_fast(start)
        if self._time_plan is not None:
            return self._load_time(start)
        dg = self._slice_op(start, start + self._batch_size)
        batch = dg.materialize()
        if self._hook_manager is not None:
            batch = self._hook_manager.execute_active_hooks(dg, batch)
        return batch

    def _load_fast(self, start: int) -> DGBatch:
        store, lo0, hi0 = self._fast
        bs = self._batch_size
        lo, hi = max(start, lo0), min(start + bs, hi0)
        if hi < lo:
            hi = lo
        n = hi - lo
        src = None
        if lo % bs == 0 and n:
            # the views of 4096 batches come from one Tensor.split per array (storage.batch_chunk)
            c, r = divmod((lo - self._origin) // bs, store._CHUNK_BATCHES)
            cur = self._chunk
            if cur is None or cur[0] != c:
                cur = self._chunk = (c, store.batch_chunk(self._origin, bs, c))
            ch = cur[1]
            src = ch[0][r]
            if src.shape[0] == n:
                dst, t = ch[1][r], ch[2][r]
                x = None if ch[3] is None else ch[3][r]
            else:
                src = None
        if src is None:
            src, dst, t = store._src[lo:hi], store._dst[lo:hi], store._t[lo:hi]
            x = None if store._x is None or not n else store._x[lo:hi]
        batch = DGBatch(src, dst, t, x)
        # which rows of the store these views are (hooks verify by identity that nobody replaced
        # the tensors since, instead of re-deriving the offsets from pointers every batch)
        if n:
            batch._slab = (store, lo, hi, src, dst, t)
        hm = self._hook_manager
        if hm is not None:
            src_dg = self._dg
            s = src_dg._slice
            view = DGraph._from_storage(store, src_dg._time_delta, src_dg._device, DGSliceTracker(
                s.start_time, s.end_time, lo, hi))
            batch = hm.execute_active_hooks(view, batch)
        return batch

    # kept for API parity with the reference, whose loader is its own collate_fn (:158)
    def __call__(self, slice_start) -> DGBatch:
        return self._load(slice_start[0])

    @staticmethod
    def _is_batch_empty(batch: DGBatch) -> bool:
        n = batch.edge_src.numel()
        n += batch.node_x_nids.numel() if batch.node_x_nids is not None else 0
        n += batch.node_y_nids.numel() if batch.node_y_nids is not None else 0
        return n == 0

    def __iter__(self) -> Iterator[DGBatch]:
        for start in self._starts:
            batch = self._load(start)
            if self._is_batch_empty(batch):
                if self._on_empty == 'raise':
                    raise EmptyBatchError('Empty batch encountered')
                if self._on_empty == 'skip':
                    continue
            yield batch
