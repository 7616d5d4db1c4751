"""Generate golden fixtures by running the UNMODIFIED reference (tgm-team/tgm @ 5183dc9)
in the build container.  Run from the repo root:

    python tests/golden/make_golden.py

Writes tests/golden/recency_*.npz.  Each file holds the inputs of one scenario and, per
loader batch and hop, what `RecencyNeighborHook` put on the batch.  The GPU box has no
/root/reference, so tests read the committed fixtures, never the reference itself.
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


class _InjectNegatives:
    """Minimal DGHook that attaches pre-generated negatives (stands in for the TGB sampler)."""
    has_state = False
    requires = {'edge_src', 'edge_dst', 'edge_time'}
    produces = {'neg', 'neg_time'}

    def __init__(self, neg: torch.Tensor, bs: int) -> None:
        self.neg, self.bs, self.i = neg, bs, 0

    def __call__(self, dg, batch):
        n = batch.edge_src.numel()
        batch.neg = self.neg[self.i:self.i + n].clone()
        batch.neg_time = batch.edge_time.clone()
        self.i += n
        return batch

    def reset_state(self) -> None:
        self.i = 0


def run_reference(src, dst, t, x, N, bs, num_nbrs, directed, neg=None, epochs=1):
    ei = torch.from_numpy(np.stack([src, dst], 1).astype(np.int32))
    data = DGData.from_raw(torch.from_numpy(t.astype(np.int64)), ei,
                           None if x is None else torch.from_numpy(x))
    dg = DGraph(data)
    keys_n, keys_t = ['edge_src', 'edge_dst'], ['edge_time', 'edge_time']
    hm = HookManager(keys=['g'])
    if neg is not None:
        hm.register('g', _InjectNegatives(torch.from_numpy(neg.astype(np.int32)), bs))
        keys_n, keys_t = keys_n + ['neg'], keys_t + ['neg_time']
    hook = RecencyNeighborHook(num_nodes=N, num_nbrs=list(num_nbrs), seed_nodes_keys=keys_n,
                               seed_times_keys=keys_t, directed=directed)
    hm.register('g', hook)
    out = {}
    with hm.activate('g'):
        for ep in range(epochs):
            for b, batch in enumerate(DGDataLoader(dg, batch_size=bs, hook_manager=hm)):
                for h in range(len(num_nbrs)):
                    tag = f'e{ep}_b{b}_h{h}'
                    out[tag + '_seed'] = batch.seed_nids[h].numpy()
                    out[tag + '_tq'] = batch.seed_times[h].numpy()
                    out[tag + '_nid'] = batch.nbr_nids[h].numpy()
                    out[tag + '_nt'] = batch.nbr_edge_time[h].numpy()
                    out[tag + '_nx'] = batch.nbr_edge_x[h].numpy()
            if ep + 1 < epochs:
                hm.reset_state()
    out['final_ids'] = hook._nbr_ids.numpy()
    out['final_times'] = hook._nbr_times.numpy()
    out['final_feats'] = hook._nbr_feats.numpy()
    out['final_write_pos'] = hook._write_pos.numpy()
    return out


def save(name, src, dst, t, x, N, bs, num_nbrs, directed, neg=None, epochs=1):
    assert N * (int(t.max()) + 1) < 2 ** 31, 'outside the parity domain (recency.py:347-348)'
    out = run_reference(src, dst, t, x, N, bs, num_nbrs, directed, neg, epochs)
    meta = dict(src=src.astype(np.int32), dst=dst.astype(np.int32), t=t.astype(np.int64),
                N=np.int64(N), bs=np.int64(bs), num_nbrs=np.array(num_nbrs, np.int64),
                directed=np.int64(directed), epochs=np.int64(epochs),
                has_x=np.int64(x is not None), has_neg=np.int64(neg is not None))
    if x is not None:
        meta['x'] = x.astype(np.float32)
    if neg is not None:
        meta['neg'] = neg.astype(np.int32)
    np.savez_compressed(os.path.join(HERE, f'recency_{name}.npz'), **meta, **out)
    print(name, 'ok', sum(v.nbytes for v in out.values()) // 1024, 'KiB raw')


def main() -> None:
    f32 = np.float32
    # --- graphs of the reference's own unit tests (test_recency_nbr_hook.py:18-49,497-520,569-590)
    s, d = np.array([0, 0, 2, 2]), np.array([1, 2, 3, 0])
    t, x = np.array([1, 2, 3, 4]), np.array([[1], [2], [5], [2]], f32)
    save('alice_1hop', s, d, t, x, 4, 1, [1], False)
    save('alice_1hop_directed', s, d, t, x, 4, 1, [1], True)
    s, d = np.zeros(100, int), np.arange(1, 101)
    save('star_exceed_buffer', s, d, np.arange(100), np.arange(1, 101, dtype=f32)[:, None], 101, 2, [2], False)
    s, d = np.array([0, 1, 3, 4, 5, 5]), np.array([1, 2, 2, 2, 0, 2])
    t, x = np.arange(1, 7), np.array([[1], [3], [5], [6], [5], [7]], f32)
    save('twohop', s, d, t, x, 6, 1, [1, 1], False)
    save('twohop_directed', s, d, t, x, 6, 1, [1, 1], True)
    save('twohop_neg', s, d, t, x, 6, 1, [1, 1], False, neg=np.array([2, 3, 0, 1, 4, 4]))
    save('twohop_nofeat', s, d, t, None, 6, 1, [1, 1], False)
    # --- seeded random streams: heavy timestamp ties, hot nodes, > B pushes per node per batch
    rng = np.random.default_rng(20261017)
    cfgs = [  # name, N, E, T, D, bs, num_nbrs, directed, neg, epochs
        ('rand_a', 10, 1200, 50, 3, 7, [2], False, False, 1),
        ('rand_b', 15, 1500, 100, 2, 50, [4, 2], False, True, 1),
        ('rand_c', 40, 1500, 300, 0, 64, [3, 5], True, False, 1),
        ('rand_d', 40, 1000, 30, 4, 64, [5, 3], False, True, 2),
        ('rand_e', 300, 1500, 1000, 1, 200, [20], False, False, 1),
        ('rand_f', 25, 900, 20, 2, 33, [2, 2, 2], False, False, 1),
    ]
    for name, N, E, T, D, bs, nn, directed, neg, epochs in cfgs:
        src = rng.integers(0, N, E)
        dst = rng.integers(0, N, E)  # self-loops allowed: they push two entries
        if name == 'rand_e':  # Zipf-ish hot node
            src = np.where(rng.random(E) < 0.3, 0, src)
        t = np.sort(rng.integers(0, T, E))
        x = rng.standard_normal((E, D)).astype(f32) if D else None
        ng = rng.integers(0, N, E) if neg else None
        save(name, src, dst, t, x, N, bs, nn, directed, ng, epochs)


if __name__ == '__main__':
    main()
