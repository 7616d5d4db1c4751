"""GPU parity of the uniform full-history sampler: the drop-in
NeighborSamplerHook / storage.get_nbrs vs fixtures from the unmodified reference, and
set-validity + uniformity where the reference falls back to CPython's random.sample."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.recency_oracle import uniform_candidates, uniform_sample_deterministic
from tests._golden import GOLDEN_DIR

pytestmark = pytest.mark.gpu

from tgm_b200 import (DGData, DGDataLoader, DGraph, HookManager,  # noqa: E402
                      NeighborSamplerHook)
from tgm_b200.core.storage import DGSliceTracker  # noqa: E402

DEV = 'cuda:0'


def _graph(src, dst, t, x):
    ei = torch.from_numpy(np.stack([src, dst], 1).astype(np.int32))
    return DGraph(DGData.from_raw(torch.from_numpy(np.asarray(t, np.int64)), ei,
                                  None if x is None else torch.from_numpy(x)), device=DEV)


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'uniform_*.npz'))),
                         ids=lambda p: os.path.basename(p)[8:-4])
def test_uniform_hook_matches_reference_fixture(path):
    z = np.load(path)
    x = z['x'] if int(z['has_x']) else None
    dg = _graph(z['src'], z['dst'], z['t'], x)
    nn = [int(v) for v in z['num_nbrs']]
    hm = HookManager(keys=['g'])
    hm.register('g', NeighborSamplerHook(num_nbrs=nn, seed_nodes_keys=['edge_src', 'edge_dst'],
                                         seed_times_keys=['edge_time', 'edge_time'],
                                         directed=bool(int(z['directed']))))
    with hm.activate('g'):
        nb = 0
        for b, batch in enumerate(DGDataLoader(dg, batch_size=int(z['bs']), hook_manager=hm)):
            nb += 1
            for h in range(len(nn)):
                for name, got in (('seed', batch.seed_nids[h]), ('tq', batch.seed_times[h]),
                                  ('nid', batch.nbr_nids[h]), ('nt', batch.nbr_edge_time[h]),
                                  ('nx', batch.nbr_edge_x[h])):
                    want = z[f'b{b}_h{h}_{name}']
                    g = got.cpu().numpy()
                    assert g.dtype == want.dtype and g.shape == want.shape, (b, h, name)
                    assert np.array_equal(g, want), (b, h, name)
    assert nb == -(-len(z['src']) // int(z['bs']))


def test_get_nbrs_subsampling_is_valid_and_uniform():
    """Seeds with more than k candidates: every returned slot is a true candidate (edge inside
    the slice, right neighbour, right time and feature row), no candidate entry appears twice,
    duplicates of a node share one draw (array_backend.py:119,166), and over many draws every
    candidate is picked about equally often."""
    rng = np.random.default_rng(5)
    N, E, D, k = 50, 4000, 3, 5
    src = rng.integers(0, N, E).astype(np.int32)
    dst = ((src + 1 + rng.integers(0, N - 1, E)) % N).astype(np.int32)  # no self-loops: a
    # self-loop contributes two identical candidate rows, which the bookkeeping below would merge
    t = np.sort(rng.integers(0, 500, E)).astype(np.int64)
    x = rng.standard_normal((E, D)).astype(np.float32)
    dg = _graph(src, dst, t, x)
    st = dg._storage
    cut_t = 300
    e_hi = int(np.searchsorted(t, cut_t, 'right'))
    seeds = np.concatenate([np.arange(N), np.arange(N), [-1]]).astype(np.int32)
    cand = uniform_candidates(src, dst, 0, e_hi, np.arange(N))
    counts = {v: {} for v in range(N)}
    draws = 400
    for r in range(draws):
        torch.manual_seed(r)
        nid, nt, nx = st.get_nbrs(torch.from_numpy(seeds), k, DGSliceTracker(end_time=cut_t), False)
        nid, nt, nx = nid.cpu().numpy(), nt.cpu().numpy(), nx.cpu().numpy()
        assert (nid[-1] == -1).all() and (nt[-1] == 0).all() and (nx[-1] == 0).all()
        assert np.array_equal(nid[:N], nid[N:2 * N]) and np.array_equal(nt[:N], nt[N:2 * N])
        if r % 50:
            for v in range(N):  # cheap bookkeeping only
                for j in range(k):
                    key = (int(nid[v, j]), int(nt[v, j]), float(nx[v, j, 0]))
                    counts[v][key] = counts[v].get(key, 0) + 1
            continue
        for v in range(N):
            rows = {(nb, int(t[e]), float(x[e, 0])): e for e, nb in cand[v]}
            assert len(cand[v]) > k
            picked = [(int(nid[v, j]), int(nt[v, j]), float(nx[v, j, 0])) for j in range(k)]
            assert all(p in rows for p in picked)
            assert len(set(picked)) == k  # feature values are unique per edge: distinct entries
            for j in range(k):
                assert np.array_equal(nx[v, j], x[rows[picked[j]]])
    # uniformity: expected count per candidate entry = draws_counted * k / len(cand)
    counted = draws - draws // 50
    worst = 0.0
    for v in range(N):
        exp = counted * k / len(cand[v])
        for c in counts[v].values():
            worst = max(worst, abs(c - exp) / np.sqrt(exp))
    assert worst < 6.0  # ~6 sigma over ~8000 cells


def test_get_nbrs_deterministic_cases_vs_oracle():
    """Mixed degrees: rows of seeds with <= k candidates equal the oracle exactly (incl. a slice
    with a start bound and directed mode); the others are checked for validity."""
    rng = np.random.default_rng(8)
    N, E, D, k = 400, 3000, 4, 8
    src, dst = rng.integers(0, N, E).astype(np.int32), rng.integers(0, N, E).astype(np.int32)
    t = np.sort(rng.integers(0, 900, E)).astype(np.int64)
    x = rng.standard_normal((E, D)).astype(np.float32)
    dg = _graph(src, dst, t, x)
    seeds = rng.integers(0, N, 1000).astype(np.int32)
    for directed in (False, True):
        for sl, (t0, t1) in ((DGSliceTracker(end_time=700), (None, 700)),
                             (DGSliceTracker(start_time=200, end_time=650), (200, 650))):
            e_lo = 0 if t0 is None else int(np.searchsorted(t, t0, 'left'))
            e_hi = int(np.searchsorted(t, t1, 'right'))
            nid, nt, nx = dg._storage.get_nbrs(torch.from_numpy(seeds), k, sl, directed)
            w_nid, w_nt, w_nx, exact = uniform_sample_deterministic(src, dst, t, x, e_lo, e_hi,
                                                                    seeds, k, directed)
            assert exact.sum() > 100 and (~exact).sum() > (3 if directed else 50)
            assert np.array_equal(nid.cpu().numpy()[exact], w_nid[exact])
            assert np.array_equal(nt.cpu().numpy()[exact], w_nt[exact])
            assert np.array_equal(nx.cpu().numpy()[exact], w_nx[exact])
            assert bool((nid[torch.from_numpy(~exact)] != -1).all())  # over-full seeds: k slots


@pytest.mark.parametrize('directed', [False, True])
def test_time_bounded_slices_equal_the_index_resolved_call(directed):
    """get_nbrs with a slice bounded by times only goes to the kernel as it is
    (tgm_csr_sample_uniform_time: candidates by entry time); the same slice resolved to edge
    indices first (tgm_store_bounds -> tgm_csr_sample_uniform) must give identical rows -- also
    beyond k candidates, the draw being keyed by (seed, node).  Ties on the bounds, empty
    intervals, unbounded sides; ids outside the graph give padding rows."""
    from tgm_b200 import _cabi
    from tgm_b200.sampler import full_history_neighbors
    rng = np.random.default_rng(17)
    N, E, T, D, k = 300, 20_000, 400, 8, 6
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    t = np.sort(rng.integers(0, T, E))
    x = rng.standard_normal((E, D)).astype(np.float32)
    dg = _graph(src, dst, t, x)
    st = dg._storage
    seeds = torch.from_numpy(np.concatenate([rng.integers(0, N, 500), [-1, N + 5]]).astype(np.int32)).to(DEV)
    for t_lo, t_hi in [(None, None), (None, 123), (57, None), (57, 57), (100, 99), (0, T), (-5, 3),
                       (T - 1, None), (None, -1)]:
        sl = DGSliceTracker(start_time=t_lo, end_time=t_hi)
        got = full_history_neighbors(st, seeds, k, sl, directed, rng_seed=99)
        lo, hi = st.edge_range(sl)
        csr = st._node_cache[('uniform_csr', directed)]
        S = seeds.numel()
        nid = torch.empty((S, k), dtype=torch.int32, device=DEV)
        nt = torch.empty((S, k), dtype=torch.int64, device=DEV)
        nx = torch.empty((S, k, D), dtype=torch.float32, device=DEV)
        _cabi.check(_cabi.lib.tgm_csr_sample_uniform(
            csr.handle, seeds.data_ptr(), S, lo, max(lo, hi), k, 99, nid.data_ptr(), nt.data_ptr(),
            nx.data_ptr(), _cabi.current_stream(torch.device(DEV))))
        for g_, w_ in zip(got, (nid, nt, nx)):
            assert torch.equal(g_, w_), (t_lo, t_hi)
        valid = got[0] != -1
        if t_lo is not None:
            assert bool((got[1][valid] >= t_lo).all())
        if t_hi is not None:
            assert bool((got[1][valid] <= t_hi).all())
