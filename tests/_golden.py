"""Helpers shared by the golden-fixture tests."""
from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, 'recency_*.npz')))


def golden_ids():
    return [os.path.basename(f)[len('recency_'):-len('.npz')] for f in golden_files()]


class Golden:
    """One fixture: inputs plus what the reference put on every batch."""

    def __init__(self, path: str) -> None:
        z = np.load(path)
        self.z = z
        self.src, self.dst, self.t = z['src'], z['dst'], z['t']
        self.x = z['x'] if int(z['has_x']) else None
        self.neg = z['neg'] if int(z['has_neg']) else None
        self.N, self.bs = int(z['N']), int(z['bs'])
        self.num_nbrs = [int(v) for v in z['num_nbrs']]
        self.directed = bool(int(z['directed']))
        self.epochs = int(z['epochs'])
        self.D = 0 if self.x is None else self.x.shape[1]
        self.E = len(self.src)

    def batches(self):
        for b, lo in enumerate(range(0, self.E, self.bs)):
            yield b, lo, min(lo + self.bs, self.E)

    def seeds(self, lo: int, hi: int):
        """hop-0 seeds in the fixture's seed_nodes_keys order: src, dst(, neg)."""
        parts = [self.src[lo:hi], self.dst[lo:hi]]
        times = [self.t[lo:hi], self.t[lo:hi]]
        if self.neg is not None:
            parts.append(self.neg[lo:hi])
            times.append(self.t[lo:hi])
        return np.concatenate(parts).astype(np.int32), np.concatenate(times).astype(np.int64)

    def expect(self, ep: int, b: int, h: int):
        tag = f'e{ep}_b{b}_h{h}'
        return tuple(self.z[f'{tag}_{n}'] for n in ('seed', 'tq', 'nid', 'nt', 'nx'))


def assert_hop_equal(got, want, where=''):
    names = ('seed_nids', 'seed_times', 'nbr_nids', 'nbr_edge_time', 'nbr_edge_x')
    for g, w, n in zip(got, want, names):
        g = np.asarray(g)
        assert g.dtype == w.dtype, f'{where} {n}: dtype {g.dtype} != {w.dtype}'
        assert g.shape == w.shape, f'{where} {n}: shape {g.shape} != {w.shape}'
        assert np.array_equal(g, w), f'{where} {n}: values differ'
