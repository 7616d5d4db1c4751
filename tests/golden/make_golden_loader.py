"""Golden fixtures for the view algebra and the loader's batch plan (SURVEY section 8a rows S1, S3,
S4), from the UNMODIFIED reference (tgm/core/graph.py:110-152, tgm/data/loader.py:101-170,
tgm/core/_storage/backends/array_backend.py:57-68, :301-321) run on CPU:

    python tests/golden/make_golden_loader.py      -> tests/golden/loader_plans.npz

Per case: a random stream (edges, optionally dynamic node features and node labels on the same
timeline), an optional chain of slice_time / slice_events views, then DGDataLoader(on_empty=None)
so that every batch -- empty ones included -- is recorded: the ids of its edge events (edge_x[:, 0]
carries the edge's index in the sorted stream), its node-event ids/times and its node-label
ids/times, plus the metadata properties of the view the loader iterates.

Parity domain.  When node events or labels are present the merged timeline is unsorted and the
reference re-orders ALL events with `torch.argsort(time)` (tgm/data/dg_data.py:351-358), which is
not a stable sort: the order among events with equal timestamps -- and with event-index batching
even which batch they fall into -- is implementation-defined.  Cases marked `exact` keep every
timestamp unique (or have edge events only, where the sorted input is left untouched) and are
compared element by element; the other cases use time-window batching, where ties cannot change a
batch's membership, and are compared per batch as sorted sets."""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shim import import_reference  # noqa: E402

import_reference()
from tgm import DGraph  # noqa: E402
from tgm.data import DGData, DGDataLoader  # noqa: E402

# name, E, N, T, n_node_events, n_labels, time_delta, view ops, loader kwargs, exact (unique times)
CASES = [
    ('events_bs7', 50, 12, 40, 0, 0, 'r', [], dict(batch_size=7), True),
    ('events_bs7_drop_last', 50, 12, 40, 0, 0, 'r', [], dict(batch_size=7, drop_last=True), True),
    ('events_exact_multiple', 60, 9, 25, 0, 0, 'r', [], dict(batch_size=10), True),
    ('events_exact_multiple_drop_last', 60, 9, 25, 0, 0, 'r', [], dict(batch_size=10, drop_last=True), True),
    ('events_with_node_events', 40, 10, 300, 15, 9, 'r', [], dict(batch_size=6), True),
    ('events_sliced_view', 80, 15, 600, 10, 0, 'r', [('slice_events', 9, 61), ('slice_time', 120, 500)],
     dict(batch_size=8), True),
    ('seconds_by_5s', 70, 11, 200, 0, 0, 's', [], dict(batch_size=5, batch_unit='s'), True),
    ('seconds_by_1m', 90, 11, 400, 12, 7, 's', [], dict(batch_size=1, batch_unit='m'), False),
    ('seconds_by_2m_drop_last', 90, 11, 400, 0, 0, 's', [], dict(batch_size=2, batch_unit='m', drop_last=True), True),
    ('seconds_sparse_with_empty_windows', 25, 8, 3000, 6, 0, 's', [], dict(batch_size=1, batch_unit='m'), True),
    ('seconds_sliced_by_30s', 120, 20, 900, 20, 10, 's', [('slice_time', 100, 700), ('slice_events', 20, 100)],
     dict(batch_size=30, batch_unit='s'), True),
    ('seconds_iterated_by_events', 64, 10, 1000, 8, 0, 's', [], dict(batch_size=16), True),
    ('heavy_ties', 100, 6, 9, 10, 10, 's', [], dict(batch_size=2, batch_unit='s'), False),
]


def build(E, N, T, n_nx, n_ny, seed, unique):
    rng = np.random.default_rng(seed)
    if unique and (n_nx or n_ny):  # every event of every kind gets its own timestamp
        pool = rng.choice(T, E + n_nx + n_ny, replace=False)
        t, nx_times, ny_times = (np.sort(p) for p in np.split(pool, [E, E + n_nx]))
    else:
        t = np.sort(rng.integers(0, T, E))
        nx_times, ny_times = np.sort(rng.integers(0, T, n_nx)), np.sort(rng.integers(0, T, n_ny))
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    kw = dict(edge_time=torch.from_numpy(t), edge_index=torch.from_numpy(np.stack([src, dst], 1)),
              edge_x=torch.from_numpy(np.stack([np.arange(E), rng.integers(0, 9, E)], 1).astype(np.float32)))
    raw = dict(t=t, src=src, dst=dst)
    if n_nx:
        raw['nx_t'], raw['nx_id'] = nx_times, rng.integers(0, N, n_nx)
        raw['nx'] = rng.standard_normal((n_nx, 3)).astype(np.float32)
        kw.update(node_x_time=torch.from_numpy(raw['nx_t']), node_x_nids=torch.from_numpy(raw['nx_id']),
                  node_x=torch.from_numpy(raw['nx']))
    if n_ny:
        raw['ny_t'], raw['ny_id'] = ny_times, rng.integers(0, N, n_ny)
        raw['ny'] = rng.standard_normal((n_ny, 2)).astype(np.float32)
        kw.update(node_y_time=torch.from_numpy(raw['ny_t']), node_y_nids=torch.from_numpy(raw['ny_id']),
                  node_y=torch.from_numpy(raw['ny']))
    return kw, raw


def main():
    out = {}
    for ci, (name, E, N, T, n_nx, n_ny, td, ops, lkw, exact) in enumerate(CASES):
        kw, raw = build(E, N, T, n_nx, n_ny, 100 + ci, exact)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            dg = DGraph(DGData.from_raw(time_delta=td, **kw))
        for op, a, b in ops:
            dg = getattr(dg, op)(a, b)
        meta = dict(start_time=dg.start_time, end_time=dg.end_time, num_events=dg.num_events,
                    num_edge_events=dg.num_edge_events, num_node_events=dg.num_node_events,
                    num_node_labels=dg.num_node_labels, num_timestamps=dg.num_timestamps,
                    num_nodes=dg.num_nodes,
                    nodes=sorted(int(v) for v in dg._storage.get_nodes(dg._slice)))
        loader = DGDataLoader(dg, on_empty=None, **lkw)
        meta['len'] = len(loader)
        eids, e_off, nx_ids, nx_t, nx_off, ny_ids, ny_t, ny_off = [], [0], [], [], [0], [], [], [0]
        for batch in loader:
            ids = batch.edge_x[:, 0].long().numpy() if batch.edge_x is not None else np.zeros(0, np.int64)
            assert len(ids) == batch.edge_src.numel()
            eids.append(ids)
            e_off.append(e_off[-1] + len(ids))
            for ids_l, t_l, off, nid, tt in ((nx_ids, nx_t, nx_off, batch.node_x_nids, batch.node_x_time),
                                             (ny_ids, ny_t, ny_off, batch.node_y_nids, batch.node_y_time)):
                n = 0 if nid is None else nid.numel()
                if n:
                    ids_l.append(nid.numpy().astype(np.int64))
                    t_l.append(tt.numpy().astype(np.int64))
                off.append(off[-1] + n)
        cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64)
        assert len(e_off) - 1 == meta['len']
        pre = f'{name}/'
        out[pre + 'meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        out[pre + 'spec'] = np.frombuffer(json.dumps(dict(time_delta=td, ops=ops, loader=lkw, exact=exact)).encode(),
                                          dtype=np.uint8)
        for k, v in raw.items():
            out[pre + 'raw_' + k] = v
        out.update({pre + 'eids': cat(eids), pre + 'e_off': np.array(e_off), pre + 'nx_ids': cat(nx_ids),
                    pre + 'nx_t': cat(nx_t), pre + 'nx_off': np.array(nx_off), pre + 'ny_ids': cat(ny_ids),
                    pre + 'ny_t': cat(ny_t), pre + 'ny_off': np.array(ny_off)})
        print(name, 'batches', meta['len'], 'empty', sum(1 for i in range(meta['len'])
              if e_off[i + 1] == e_off[i] and nx_off[i + 1] == nx_off[i] and ny_off[i + 1] == ny_off[i]))
    np.savez_compressed(os.path.join(HERE, 'loader_plans.npz'), **out)


if __name__ == '__main__':
    main()
