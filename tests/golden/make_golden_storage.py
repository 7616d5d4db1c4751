"""The storage contract, getter by getter, from the UNMODIFIED reference backend
(tgm/core/_storage/backends/array_backend.py: `DGStorageArrayBackend`, the only backend the
reference registers) run on CPU in the build container:

    python tests/golden/make_golden_storage.py      -> tests/golden/storage_contract.npz

Per case (edge-only stream; edges + dynamic node features + node labels on one timeline with
unique timestamps; edge types + static node features) and per slice (time bounds, event-index
bounds, both, empty): what every `DGStorageBase` getter (tgm/core/_storage/base.py:20-118) returns.
`get_nbrs` is recorded with k >= every degree (no `random.sample` involved), undirected and
directed.  tests/test_storage_contract.py replays the same calls on `DeviceCOOStorage`."""
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
from tgm.core._storage.backends.array_backend import DGStorageArrayBackend  # noqa: E402
from tgm.core._storage.base import DGSliceTracker  # noqa: E402
from tgm.data import DGData  # noqa: E402

SLICES = [  # (start_time, end_time, start_idx, end_idx)
    (None, None, None, None), (None, 150, None, None), (60, None, None, None), (60, 150, None, None),
    (None, None, 5, 31), (40, 260, 3, 40), (100, 100, None, None), (299, 1000, None, None),
    (None, None, 20, 20), (0, 0, None, None), (500, 600, None, None), (None, None, 0, 1)]


def cases():
    rng = np.random.default_rng(7)
    E, N, T = 48, 14, 300
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    x = rng.standard_normal((E, 3)).astype(np.float32)
    t_ties = np.sort(rng.integers(0, T, E))
    yield 'edges_only', dict(edge_time=t_ties, edge_index=np.stack([src, dst], 1), edge_x=x)
    yield 'no_features', dict(edge_time=t_ties, edge_index=np.stack([src, dst], 1))
    pool = rng.choice(T, E + 10 + 7, replace=False)  # unique timestamps over all event kinds
    t, nxt, nyt = (np.sort(p) for p in np.split(pool, [E, E + 10]))
    yield 'with_node_events', dict(
        edge_time=t, edge_index=np.stack([src, dst], 1), edge_x=x,
        node_x_time=nxt, node_x_nids=rng.integers(0, N, 10), node_x=rng.standard_normal((10, 2)).astype(np.float32),
        node_y_time=nyt, node_y_nids=rng.integers(0, N, 7), node_y=rng.standard_normal((7, 4)).astype(np.float32))
    yield 'typed_static', dict(
        edge_time=t_ties, edge_index=np.stack([src, dst], 1), edge_x=x,
        edge_type=rng.integers(0, 3, E).astype(np.int32), node_type=rng.integers(0, 2, N).astype(np.int32),
        static_node_x=rng.standard_normal((N, 5)).astype(np.float32))


def sparse_parts(sp):
    if sp is None:
        return None
    sp = sp.coalesce()
    return dict(idx=sp.indices().numpy(), val=sp.values().numpy(), shape=np.array(sp.shape))


def main():
    out = {}
    for name, kw in cases():
        tk = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in kw.items()}
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            data = DGData.from_raw(time_delta='s', **tk)
        st = DGStorageArrayBackend(data)
        for k, v in kw.items():
            out[f'{name}/in_{k}'] = v
        out[f'{name}/dims'] = np.frombuffer(json.dumps(dict(
            static_node_x_dim=st.get_static_node_x_dim(), node_x_dim=st.get_node_x_dim(),
            node_y_dim=st.get_node_y_dim(), edge_x_dim=st.get_edge_x_dim(),
            has_static=st.get_static_node_x() is not None, has_node_type=st.get_node_type() is not None,
        )).encode(), dtype=np.uint8)
        for si, (t0, t1, i0, i1) in enumerate(SLICES):
            s = DGSliceTracker(t0, t1, i0, i1)
            pre = f'{name}/s{si}/'
            meta = dict(slice=[t0, t1, i0, i1], start_time=st.get_start_time(s), end_time=st.get_end_time(s),
                        nodes=sorted(int(v) for v in st.get_nodes(s)), num_events=st.get_num_events(s),
                        num_timestamps=st.get_num_timestamps(s))
            e_src, e_dst, e_t = st.get_edges(s)
            assert e_src.dtype == torch.int32 and e_t.dtype == torch.int64
            out[pre + 'edges'] = np.stack([e_src.numpy(), e_dst.numpy(), e_t.numpy()])
            ex = st.get_edge_x(s)
            meta['edge_x_none'] = ex is None
            if ex is not None:
                out[pre + 'edge_x'] = ex.numpy()
            et = st.get_edge_type(s)
            meta['edge_type_none'] = et is None
            if et is not None:
                out[pre + 'edge_type'] = et.numpy()
            for tag, (ids, tt) in (('node_events', st.get_node_events(s)), ('node_labels', st.get_node_labels(s))):
                out[pre + tag] = np.stack([ids.numpy().astype(np.int64), tt.numpy().astype(np.int64)])
            for tag, sp in (('node_x', st.get_node_x(s)), ('node_y', st.get_node_y(s))):
                parts = sparse_parts(sp)
                meta[tag + '_none'] = parts is None
                if parts is not None:
                    for k_, v in parts.items():
                        out[pre + f'{tag}_{k_}'] = v
            # get_nbrs over the slice: every seed keeps all its candidates (k = 64).  Only on
            # edge-only timelines: with node events the reference indexes edge_x by EVENT index
            # (array_backend.py:160) and raises IndexError
            if len(e_src) and 'node_x_time' not in kw:
                seeds = torch.unique(torch.cat([e_src, e_dst]))[:6]
                out[pre + 'nbr_seeds'] = seeds.numpy()
                for directed in (False, True):
                    nid, nt, nx = st.get_nbrs(seeds, 64, s, directed)
                    tag = 'nbrs_dir' if directed else 'nbrs'
                    out[pre + tag + '_nid'], out[pre + tag + '_t'] = nid.numpy(), nt.numpy()
                    out[pre + tag + '_x'] = nx.numpy()
            out[pre + 'meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        print(name, 'ok')
    np.savez_compressed(os.path.join(HERE, 'storage_contract.npz'), **out)


if __name__ == '__main__':
    main()
