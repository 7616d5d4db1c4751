"""Golden fixtures for the uniform sampler, from the UNMODIFIED reference
(NeighborSamplerHook -> DGStorageArrayBackend.get_nbrs):  python tests/golden/make_golden_uniform.py

Graphs are the reference's own unit-test graphs (test/unit/test_hooks/test_neighbor_sampler_hook.py)
plus seeded random streams whose k is >= every node's degree, so no `random.sample` is involved and
the reference output is deterministic (tests/golden/uniform_*.npz), plus streams with k below the
degrees under a fixed `random.seed` (tests/golden/uniformrng_*.npz)."""
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
from tgm.hooks import HookManager, NeighborSamplerHook  # noqa: E402


def save(name, src, dst, t, x, bs, num_nbrs, directed, rng_seed=None):
    ei = torch.from_numpy(np.stack([src, dst], 1).astype(np.int32))
    data = DGData.from_raw(torch.from_numpy(t.astype(np.int64)), ei,
                           None if x is None else torch.from_numpy(x))
    dg = DGraph(data)
    hm = HookManager(keys=['g'])
    hm.register('g', NeighborSamplerHook(num_nbrs=list(num_nbrs),
                                         seed_nodes_keys=['edge_src', 'edge_dst'],
                                         seed_times_keys=['edge_time', 'edge_time'],
                                         directed=directed))
    out = {}
    if rng_seed is not None:  # get_nbrs sub-samples with CPython's global generator (:152-153)
        import random
        random.seed(rng_seed)
    with hm.activate('g'):
        for b, batch in enumerate(DGDataLoader(dg, batch_size=bs, hook_manager=hm)):
            for h in range(len(num_nbrs)):
                tag = f'b{b}_h{h}'
                out[tag + '_seed'] = batch.seed_nids[h].numpy()
                out[tag + '_tq'] = batch.seed_times[h].numpy()
                out[tag + '_nid'] = batch.nbr_nids[h].numpy()
                out[tag + '_nt'] = batch.nbr_edge_time[h].numpy()
                out[tag + '_nx'] = batch.nbr_edge_x[h].numpy()
    meta = dict(src=src.astype(np.int32), dst=dst.astype(np.int32), t=t.astype(np.int64),
                bs=np.int64(bs), num_nbrs=np.array(num_nbrs, np.int64),
                directed=np.int64(directed), has_x=np.int64(x is not None),
                rng_seed=np.int64(-1 if rng_seed is None else rng_seed))
    if x is not None:
        meta['x'] = x.astype(np.float32)
    prefix = 'uniform_' if rng_seed is None else 'uniformrng_'
    np.savez_compressed(os.path.join(HERE, f'{prefix}{name}.npz'), **meta, **out)
    print(name, 'ok')


def main():
    f32 = np.float32
    s, d = np.array([0, 0, 2, 2]), np.array([1, 2, 3, 0])
    t, x = np.array([1, 2, 3, 4]), np.array([[1], [2], [5], [2]], f32)
    save('alice', s, d, t, x, 1, [2], False)
    save('alice_directed', s, d, t, x, 1, [2], True)
    s, d = np.array([0, 1, 3, 4, 5, 5]), np.array([1, 2, 2, 2, 0, 2])
    t, x = np.arange(1, 7), np.array([[1], [3], [5], [6], [5], [7]], f32)
    save('twohop', s, d, t, x, 1, [4, 4], False)
    save('twohop_nofeat', s, d, t, None, 1, [4, 4], False)
    rng = np.random.default_rng(99)
    for name, N, E, T, D, bs, nn, directed in [('rand_a', 40, 120, 60, 3, 10, [16], False),
                                               ('rand_b', 30, 90, 40, 2, 7, [14, 14], False),
                                               ('rand_c', 25, 100, 30, 0, 9, [18], True)]:
        src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
        t = np.sort(rng.integers(0, T, E))
        x = rng.standard_normal((E, D)).astype(f32) if D else None
        deg = np.bincount(np.concatenate([src, dst]), minlength=N).max()
        assert deg <= min(nn), (name, deg)  # k >= every degree: no random.sample on any query
        save(name, src, dst, t, x, bs, nn, directed)
    # sub-sampled queries: k below the degrees, random.seed() fixed before the loader loop, so the
    # reference's random.sample stream -- one draw per unique seed node with more than k
    # candidates, in ascending node order, hop after hop, batch after batch -- is reproducible
    for name, N, E, T, D, bs, nn, directed, seed in [('sub_a', 12, 150, 60, 3, 10, [4], False, 7),
                                                     ('sub_b', 10, 120, 40, 2, 8, [3, 2], False, 8),
                                                     ('sub_c', 8, 100, 30, 0, 9, [5], True, 9)]:
        src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
        t = np.sort(rng.integers(0, T, E))
        x = rng.standard_normal((E, D)).astype(f32) if D else None
        save(name, src, dst, t, x, bs, nn, directed, rng_seed=seed)


if __name__ == '__main__':
    main()
