"""Fixtures for time-window batching through the recency hook, from the UNMODIFIED reference:

    python tests/golden/make_golden_timeunit.py      # build container only

DGDataLoader(dg, batch_size=bs, batch_unit='s') + RecencyNeighborHook (tgm/data/loader.py:101-156,
tgm/hooks/neighbors/recency.py:119-171) over small streams with uneven, partly empty time windows.
Writes tests/golden/timeunit_*.npz: per yielded batch its edge range and what the hook put on it.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shim import import_reference  # noqa: E402

import_reference()
from tgm import DGraph  # noqa: E402
from tgm.data import DGData, DGDataLoader  # noqa: E402
from tgm.hooks import HookManager, RecencyNeighborHook  # noqa: E402


def save(name, src, dst, t, x, N, bs, num_nbrs, directed):
    assert N * (int(t.max()) + 1) < 2 ** 31
    ei = torch.from_numpy(np.stack([src, dst], 1).astype(np.int32))
    data = DGData.from_raw(torch.from_numpy(t.astype(np.int64)), ei,
                           None if x is None else torch.from_numpy(x), time_delta='s')
    dg = DGraph(data)
    hm = HookManager(keys=['g'])
    hm.register('g', RecencyNeighborHook(num_nodes=N, num_nbrs=list(num_nbrs),
                                         seed_nodes_keys=['edge_src', 'edge_dst'],
                                         seed_times_keys=['edge_time', 'edge_time'],
                                         directed=directed))
    out, seen, nb = {}, 0, 0
    with hm.activate('g'):
        for batch in DGDataLoader(dg, batch_size=bs, batch_unit='s', hook_manager=hm):
            n = batch.edge_src.numel()
            assert n > 0 and batch.edge_time.tolist() == t[seen:seen + n].tolist()
            out[f'b{nb}_lo'], out[f'b{nb}_hi'] = np.int64(seen), np.int64(seen + n)
            for h in range(len(num_nbrs)):
                tag = f'b{nb}_h{h}'
                out[tag + '_seed'] = batch.seed_nids[h].numpy()
                out[tag + '_tq'] = batch.seed_times[h].numpy()
                out[tag + '_nid'] = batch.nbr_nids[h].numpy()
                out[tag + '_nt'] = batch.nbr_edge_time[h].numpy()
                out[tag + '_nx'] = batch.nbr_edge_x[h].numpy()
            seen += n
            nb += 1
    assert seen == len(src)
    meta = dict(src=src.astype(np.int32), dst=dst.astype(np.int32), t=t.astype(np.int64),
                N=np.int64(N), bs=np.int64(bs), num_nbrs=np.array(num_nbrs, np.int64),
                directed=np.int64(directed), nb=np.int64(nb), has_x=np.int64(x is not None))
    if x is not None:
        meta['x'] = x.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, f'timeunit_{name}.npz'), **meta, **out)
    print(name, nb, 'batches')


def main():
    rng = np.random.default_rng(20261018)
    for name, N, E, T, D, bs, nn, directed, gaps in [
            ('a', 30, 900, 400, 3, 7, [4], False, False),
            ('b', 20, 700, 3000, 2, 11, [3, 2], False, True),     # many empty windows, 2 hops
            ('c', 25, 800, 90, 0, 5, [5], True, False)]:          # heavy ties, directed, no feats
        src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
        t = np.sort(rng.integers(0, T, E))
        if gaps:
            t = np.sort(np.where(rng.random(E) < 0.5, t // 40 * 40, t))
        x = rng.standard_normal((E, D)).astype(np.float32) if D else None
        save(name, src, dst, t, x, N, bs, nn, directed)


if __name__ == '__main__':
    main()
