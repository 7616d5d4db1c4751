"""DGData: validated, time-sorted host container the device store ingests.

Mirrors the fields and `from_raw` signature of tgm/data/dg_data.py:29-84,591-674 and the
validation that shapes the hot path (:86-394): timestamps int64, non-negative, < int32 max
(:124-136); ids int32 (:161); PADDED_NODE_ID rejected; events globally time-sorted, sorting
everything if they are not (:350-394).  One-time host ingest -- out of the hot path (SURVEY.md
section 2): splits, discretisation and the CSV/pandas/TGB constructors are not rebuilt here.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from tgm_b200.constants import PADDED_NODE_ID
from tgm_b200.core.timedelta import TimeDeltaDG
from tgm_b200.exceptions import EmptyGraphError, InvalidNodeIDError

_INT_TYPES = (torch.int8, torch.int16, torch.int32, torch.int64, torch.uint8)
_INT32_MAX = torch.iinfo(torch.int32).max


def _as_tensor(x, name: str) -> Tensor:
    if not isinstance(x, Tensor):
        raise TypeError(f'{name} must be a Tensor, got: {type(x)}')
    if x.is_floating_point() and torch.isnan(x).any():
        raise ValueError(f'{name} contains NaN values')
    return x


def _integral(x, name: str) -> Tensor:
    x = _as_tensor(x, name)
    if x.dtype not in _INT_TYPES:
        raise TypeError(f'{name} must have integer dtype but got: {x.dtype}')
    return x


def _to_int32(x: Tensor, name: str) -> Tensor:
    if x.dtype == torch.int64:
        warnings.warn(f'Downcasting {name} from torch.int64 to torch.int32', UserWarning)
    return x.to(torch.int32)


def _ids(x, name: str) -> Tensor:
    """Node ids: integral, not the padded id, inside the int32 range (dg_data.py:148-161,
    :199-213); negative ids are rejected here as well (they can only corrupt the samplers)."""
    x = _integral(x, name)
    if x.numel() and (x == PADDED_NODE_ID).any():
        raise InvalidNodeIDError(
            f'{name} contains node ids matching PADDED_NODE_ID: {PADDED_NODE_ID}, which is used '
            'to mark invalid neighbors. Try remapping all node ids to positive integers.')
    if x.numel() and (x < 0).any():
        raise InvalidNodeIDError(f'{name} contains negative node ids')
    if x.numel() and (x >= _INT32_MAX).any():
        raise InvalidNodeIDError(f'{name} contains node ids that exceed the int32 limit '
                                 f'({_INT32_MAX})')
    return _to_int32(x, name)


def _feats(x, name: str, rows: Optional[int], what: str = '') -> Tensor:
    x = _as_tensor(x, name)
    if x.ndim != 2 or (rows is not None and x.shape[0] != rows):
        raise ValueError(f'{name} must have shape [{rows if rows is not None else "N"}, D]'
                         f'{what}, got {tuple(x.shape)}')
    if x.dtype == torch.float64:
        warnings.warn(f'Downcasting {name} from torch.float64 to torch.float32', UserWarning)
    return x.to(torch.float32)


def _mask(x, name: str) -> Tensor:
    """Event positions inside the timeline; int32 like upstream (the event count fits)."""
    return _integral(x, name).to(torch.int32)


@dataclass
class DGData:
    time_delta: 'TimeDeltaDG | str'
    time: Tensor                      # int64 [num_events], non-decreasing
    edge_mask: Tensor                 # position of every edge event inside `time`, ascending
    edge_index: Tensor                # int32 [E, 2]
    edge_x: Optional[Tensor] = None   # float32 [E, D]
    node_x_mask: Optional[Tensor] = None
    node_x_nids: Optional[Tensor] = None
    node_x: Optional[Tensor] = None
    node_y_mask: Optional[Tensor] = None
    node_y_nids: Optional[Tensor] = None
    node_y: Optional[Tensor] = None
    static_node_x: Optional[Tensor] = None
    edge_type: Optional[Tensor] = None
    node_type: Optional[Tensor] = None

    def __post_init__(self) -> None:
        if isinstance(self.time_delta, str):
            self.time_delta = TimeDeltaDG(self.time_delta)
        t = _integral(self.time, 'timestamps')
        if t.numel() == 0:
            raise EmptyGraphError('Cannot construct a graph without events')
        if (t < 0).any():
            raise ValueError('timestamps must all be non-negative')
        if (t >= _INT32_MAX).any():
            raise ValueError(f'timestamps exceed the int32 limit ({_INT32_MAX})')
        self.time = t.to(torch.int64)

        ei = _integral(self.edge_index, 'edge_index')
        if ei.ndim != 2 or ei.shape[1] != 2:
            raise ValueError(f'edge_index must have shape [num_edges, 2], got {tuple(ei.shape)}')
        self.edge_index = _ids(ei, 'edge_index')
        E = ei.shape[0]
        if E == 0:
            raise EmptyGraphError('TGM does not support graphs without edge events')
        self.edge_mask = _mask(self.edge_mask, 'edge_mask')
        if self.edge_mask.shape != (E,):
            raise ValueError('edge_mask must have one entry per edge')
        if self.edge_x is not None:
            self.edge_x = _feats(self.edge_x, 'edge_x', E)

        def node_events(mask, nids, feats, kind):
            """(mask, ids, features, count) of the node events / node labels (:186-263)."""
            mask = _mask(mask, f'{kind}_mask')
            n = mask.shape[0]
            if n == 0:
                raise ValueError(f'{kind}_mask is an empty tensor, please double-check your inputs')
            nids = _integral(nids, f'{kind}_nids')
            if nids.ndim != 1 or nids.shape[0] != n:
                raise ValueError(f'{kind}_nids must have shape [{n}], got {tuple(nids.shape)}')
            nids = _ids(nids, f'{kind}_nids')
            if feats is not None:
                feats = _feats(feats, kind, n)
            return mask, nids, feats, n

        n_nx = n_ny = 0
        if self.node_x_mask is not None:
            self.node_x_mask, self.node_x_nids, self.node_x, n_nx = node_events(
                self.node_x_mask, self.node_x_nids, self.node_x, 'node_x')
        if self.node_y_mask is not None:
            self.node_y_mask, self.node_y_nids, self.node_y, n_ny = node_events(
                self.node_y_mask, self.node_y_nids, self.node_y, 'node_y')

        num_nodes = int(self.edge_index.max()) + 1
        if n_nx:
            num_nodes = max(num_nodes, int(self.node_x_nids.max()) + 1)
        if n_ny and int(self.node_y_nids.max()) >= num_nodes:
            raise InvalidNodeIDError(
                "Dynamic node labels (node_y) reference node IDs outside the graph's node ID "
                f'range ({num_nodes} nodes)')
        if self.static_node_x is not None:
            sx = _feats(self.static_node_x, 'static_node_x', None)
            if sx.shape[0] < num_nodes:
                raise ValueError(f'static_node_x has shape {tuple(sx.shape)} but the data requires '
                                 f'features for at least {num_nodes} nodes')
            self.static_node_x = sx
        if self.edge_type is not None:
            # upstream validates and warns but keeps the dtype (:321-329)
            et = _integral(self.edge_type, 'edge_type')
            if et.ndim != 1 or et.shape[0] != E:
                raise ValueError(f'edge_type must have shape [num_edges], got {E} edges and '
                                 f'shape {tuple(et.shape)}')
            _to_int32(et, 'edge_type')
        if self.node_type is not None:
            nt = _integral(self.node_type, 'node_type')
            if nt.ndim != 1 or nt.shape[0] < num_nodes:
                raise ValueError(f'node_type must have shape [num_nodes], got {num_nodes} nodes '
                                 f'and shape {tuple(nt.shape)}')
            _to_int32(nt, 'node_type')
        self._num_nodes = num_nodes

        if self.time.ndim != 1 or self.time.shape[0] != E + n_nx + n_ny:
            raise ValueError(
                'time must have shape [num_edges + num_node_events + num_node_labels], got '
                f'{E} edges, {n_nx} node events, {n_ny} node labels, shape {tuple(self.time.shape)}')

        if (self.time[1:] < self.time[:-1]).any():
            self._sort_events()

    def _sort_events(self) -> None:
        """Globally time-sort all events and permute every per-event array (dg_data.py:350-394).
        The sort is stable, so simultaneous events keep their given order."""
        warnings.warn('Timestamps are not globally sorted: reordering all events', UserWarning)
        order = torch.argsort(self.time, stable=True)
        rank = torch.empty_like(order)
        rank[order] = torch.arange(order.numel())
        rank = rank.to(torch.int32)
        self.time = self.time[order]

        def reorder(mask, *arrays):
            new_pos = rank[mask.long()]
            perm = torch.argsort(new_pos, stable=True)
            return (new_pos[perm], *[None if a is None else a[perm] for a in arrays])

        self.edge_mask, self.edge_index, self.edge_x, self.edge_type = reorder(
            self.edge_mask, self.edge_index, self.edge_x, self.edge_type)
        if self.node_x_mask is not None:
            self.node_x_mask, self.node_x_nids, self.node_x = reorder(
                self.node_x_mask, self.node_x_nids, self.node_x)
        if self.node_y_mask is not None:
            self.node_y_mask, self.node_y_nids, self.node_y = reorder(
                self.node_y_mask, self.node_y_nids, self.node_y)

    @property
    def num_nodes(self) -> int:
        return self._num_nodes

    @property
    def num_edges(self) -> int:
        return self.edge_index.shape[0]

    @classmethod
    def from_raw(cls, edge_time: Tensor, edge_index: Tensor, edge_x: Tensor | None = None,
                 node_x_time: Tensor | None = None, node_x_nids: Tensor | None = None,
                 node_x: Tensor | None = None, node_y_time: Tensor | None = None,
                 node_y_nids: Tensor | None = None, node_y: Tensor | None = None,
                 static_node_x: Tensor | None = None, time_delta: 'TimeDeltaDG | str' = 'r',
                 edge_type: Tensor | None = None, node_type: Tensor | None = None) -> 'DGData':
        """Same arguments as tgm.data.DGData.from_raw: one timeline = edge times, then node-event
        times, then node-label times; the masks are each event's position in it."""
        edge_time = _as_tensor(edge_time, 'edge_time')
        parts, E = [edge_time], edge_time.shape[0]
        nx_mask = ny_mask = None
        if node_x_time is not None:
            nx_mask = torch.arange(E, E + node_x_time.shape[0])
            parts.append(_as_tensor(node_x_time, 'node_x_time'))  # cat promotes: floats are refused below
        if node_y_time is not None:
            off = E + (0 if node_x_time is None else node_x_time.shape[0])
            ny_mask = torch.arange(off, off + node_y_time.shape[0])
            parts.append(_as_tensor(node_y_time, 'node_y_time'))
        return cls(time_delta=time_delta, time=torch.cat(parts), edge_mask=torch.arange(E),
                   edge_index=edge_index, edge_x=edge_x, node_x_mask=nx_mask,
                   node_x_nids=node_x_nids, node_x=node_x, node_y_mask=ny_mask,
                   node_y_nids=node_y_nids, node_y=node_y, static_node_x=static_node_x,
                   edge_type=edge_type, node_type=node_type)
