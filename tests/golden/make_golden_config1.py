"""Config 1 pinned on the UNMODIFIED reference, negatives included.

    python tests/golden/make_golden_config1.py         # from the repo root, build container only

Runs tgm-team/tgm @ 5183dc9 (DGDataLoader(batch_size=200) + RandomNegativeEdgeSamplerHook +
RecencyNeighborHook(num_nbrs=[10], seeds src + dst + neg) on CPU) over one epoch of the
wiki-shaped synthetic stream (N=9,227 bipartite 8,227 x 1,000, E=157,474, D=172, t < 2,678,374 --
tgbl-wiki's dimensions; the dataset itself is not available offline) and writes
tests/golden/config1_wiki_epoch.npz:

  * the negatives the reference's own sampler drew (torch.manual_seed(1337), low/high = the
    destination id range, tgm/hooks/recipe.py:62-67),
  * per loader batch the position-sensitive checksums (oracle/recency_ring.c's) of nbr_nids,
    nbr_edge_time and nbr_edge_x, and the full ids/times of every 97th batch,
  * checksums of the generated inputs, so a test that regenerates the stream from the numpy seed
    knows it is looking at the same stream.

This stream lies OUTSIDE the reference's int32 sort-key domain (N * (t_max + 1) >= 2^31,
tgm/hooks/neighbors/recency.py:347-348), where its output is formally undefined; the
script therefore also drives the C oracle (ideal semantics) over the same seeds and records which
batches, if any, the unmodified reference answers differently (`differs_from_ideal`).  For those
batches `patched_csum` holds the answer of the reference with its one-token fix
(`node_ids * max_time` -> `node_ids.long() * max_time`, applied to a copy of `_update` compiled from
`inspect.getsource`, never to /root/reference), which the script checks equals the ideal semantics
on every batch of the epoch.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from _ref_shim import import_reference  # noqa: E402

import_reference()
from tgm import DGraph  # noqa: E402
from tgm.data import DGData, DGDataLoader  # noqa: E402
from tgm.hooks import HookManager, RandomNegativeEdgeSamplerHook, RecencyNeighborHook  # noqa: E402

from oracle.c_oracle import CRing, checksum_np  # noqa: E402

SEED, E, N, D, BS, K = 1, 157_474, 9227, 172, 200, 10
N_USERS = 8227


def wiki_shaped_stream():
    rng = np.random.default_rng(SEED)
    src = rng.integers(0, N_USERS, E).astype(np.int32)
    dst = rng.integers(N_USERS, N, E).astype(np.int32)
    t = np.sort(rng.integers(0, 2_678_374, E)).astype(np.int64)
    x = rng.standard_normal((E, D)).astype(np.float32)
    return src, dst, t, x


def long_key_hook_class():
    """RecencyNeighborHook with the int32 sort key widened."""
    import inspect
    import textwrap

    import tgm.hooks.neighbors.recency as mod
    code = textwrap.dedent(inspect.getsource(mod.RecencyNeighborHook._update))
    assert 'node_ids * max_time' in code
    ns = dict(vars(mod))
    exec(code.replace('node_ids * max_time', 'node_ids.long() * max_time'), ns)

    class Patched(mod.RecencyNeighborHook):
        _cls_requires = mod.RecencyNeighborHook._cls_requires
        _cls_produces = mod.RecencyNeighborHook._cls_produces
        _update = ns['_update']
    return Patched


def run(hook_cls, src, dst, t, x):
    data = DGData.from_raw(torch.from_numpy(t), torch.from_numpy(np.stack([src, dst], 1)),
                           torch.from_numpy(x))
    dg = DGraph(data)
    hm = HookManager(keys=['train'])
    hm.register('train', RandomNegativeEdgeSamplerHook(low=int(dst.min()), high=int(dst.max())))
    hm.register('train', hook_cls(
        num_nodes=N, num_nbrs=[K], seed_nodes_keys=['edge_src', 'edge_dst', 'neg'],
        seed_times_keys=['edge_time', 'edge_time', 'neg_time']))
    return dg, hm


def main() -> None:
    src, dst, t, x = wiki_shaped_stream()
    dg, hm = run(RecencyNeighborHook, src, dst, t, x)
    oracle = CRing(N, [K], D)
    torch.manual_seed(1337)  # examples/linkproppred/tgat.py:22,135
    neg, csum, keep, differs = [], [], {}, []
    with hm.activate('train'):
        for b, batch in enumerate(DGDataLoader(dg, batch_size=BS, hook_manager=hm)):
            lo, hi = b * BS, min((b + 1) * BS, E)
            assert torch.equal(batch.neg_time, batch.edge_time)
            ng = batch.neg.numpy()
            neg.append(ng)
            nid, nt, nx = (batch.nbr_nids[0].numpy(), batch.nbr_edge_time[0].numpy(),
                           batch.nbr_edge_x[0].numpy())
            assert np.array_equal(batch.seed_nids[0].numpy(), np.concatenate([src[lo:hi], dst[lo:hi], ng]))
            csum.append([checksum_np(nid), checksum_np(nt), checksum_np(nx)])
            if b % 97 == 0 or hi == E:
                keep[f'b{b}_nid'], keep[f'b{b}_nt'] = nid, nt
            seeds = np.concatenate([src[lo:hi], dst[lo:hi], ng]).astype(np.int32)
            tq = np.concatenate([t[lo:hi]] * 3)
            w = oracle.hook_call(seeds, tq, src[lo:hi], dst[lo:hi], t[lo:hi], x[lo:hi])[0]
            if not (np.array_equal(w[2], nid) and np.array_equal(w[3], nt) and np.array_equal(w[4], nx)):
                differs.append(b)
    nb = len(csum)
    print(f'{nb} batches; unmodified reference differs from the ideal semantics on {len(differs)}: {differs[:20]}')
    # the reference with the one-token fix: must equal the ideal semantics on EVERY batch
    dg2, hm2 = run(long_key_hook_class(), src, dst, t, x)
    oracle2 = CRing(N, [K], D)
    torch.manual_seed(1337)
    patched = {}
    all_neg = np.concatenate(neg)
    with hm2.activate('train'):
        for b, batch in enumerate(DGDataLoader(dg2, batch_size=BS, hook_manager=hm2)):
            lo, hi = b * BS, min((b + 1) * BS, E)
            ng = batch.neg.numpy()
            assert np.array_equal(ng, all_neg[lo:hi])
            seeds = np.concatenate([src[lo:hi], dst[lo:hi], ng]).astype(np.int32)
            w = oracle2.hook_call(seeds, np.concatenate([t[lo:hi]] * 3), src[lo:hi], dst[lo:hi],
                                  t[lo:hi], x[lo:hi])[0]
            got = (batch.nbr_nids[0].numpy(), batch.nbr_edge_time[0].numpy(),
                   batch.nbr_edge_x[0].numpy())
            assert all(np.array_equal(a, b_) for a, b_ in zip(w[2:], got)), f'patched != ideal, batch {b}'
            if b in differs:
                patched[b] = [checksum_np(v) for v in got]
    print('reference with node_ids.long() == ideal semantics on all', nb, 'batches')
    np.savez_compressed(
        os.path.join(HERE, 'config1_wiki_epoch.npz'),
        seed=np.int64(SEED), E=np.int64(E), N=np.int64(N), D=np.int64(D), bs=np.int64(BS),
        k=np.int64(K), neg=np.concatenate(neg).astype(np.int32),
        csum=np.array(csum, dtype=np.uint64), differs_from_ideal=np.array(differs, np.int64),
        patched_csum=np.array([patched[b] for b in differs], dtype=np.uint64).reshape(-1, 3),
        input_csum=np.array([checksum_np(src), checksum_np(dst), checksum_np(t), checksum_np(x)],
                            dtype=np.uint64), **keep)


if __name__ == '__main__':
    main()
