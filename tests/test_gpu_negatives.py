"""Negatives as seed producers inside the pre-sampled window:
tgm_negatives_window draws, in one launch, exactly what consecutive per-batch
torch.randint(low, high, (n,), dtype=int32, device='cuda') calls draw -- the stream of the
reference's RandomNegativeEdgeSamplerHook in device='cuda' mode
(tgm/hooks/negatives/sampler.py:45-65) -- and the default-constructed hooks serve
[src | dst | neg] neighbourhoods of a whole window from one launch per hop, with the same batch
contents as the batch-by-batch path and as the reference fixtures."""
import numpy as np
import pytest
import torch

from tests._golden import Golden, assert_hop_equal, golden_files, golden_ids

pytestmark = pytest.mark.gpu

from tgm_b200 import (DGData, DGDataLoader, DGraph, HookManager,  # noqa: E402
                      RandomNegativeEdgeSamplerHook, RecencyNeighborHook, _cabi)
from tgm_b200.hooks.negatives import SeedWindow  # noqa: E402

DEV = 'cuda:0'


def _gen():
    return torch.cuda.default_generators[0]


@pytest.mark.parametrize('n,batches,low,high', [
    (200, 7, 0, 1000), (1, 5, 3, 4), (256, 3, 8227, 9227), (257, 4, 0, 1 << 20),
    (1000, 3, -5, 5), (4096, 2, 0, 2 ** 28 - 1), (200, 1030, 0, 9227)])
def test_window_draw_equals_consecutive_torch_randint_calls(n, batches, low, high):
    torch.manual_seed(1337)
    torch.rand(5, device=DEV)  # some earlier consumer: the offset is not 0
    seed, off = _gen().initial_seed(), _gen().get_offset()
    want = torch.cat([torch.randint(low, high, (n,), dtype=torch.int32, device=DEV)
                      for _ in range(batches)])
    off_after = _gen().get_offset()
    assert off_after == off + 4 * batches, 'ATen advances the Philox offset by 4 per call'
    total = n * batches - n // 3  # the last batch is short: its call draws fewer elements
    got = torch.empty((total,), dtype=torch.int32, device=DEV)
    _cabi.check(_cabi.lib.tgm_negatives_window(seed, off, low, high, n, total, got.data_ptr(),
                                               torch.cuda.current_stream(DEV).cuda_stream))
    full = n * (batches - 1)
    assert torch.equal(got[:full], want[:full])
    torch.manual_seed(1337)
    torch.rand(5, device=DEV)
    for _ in range(batches - 1):
        torch.randint(low, high, (n,), dtype=torch.int32, device=DEV)
    last = torch.randint(low, high, (n - n // 3,), dtype=torch.int32, device=DEV)
    assert torch.equal(got[full:], last)


def test_window_draw_argument_errors():
    out = torch.empty(10, dtype=torch.int32, device=DEV)
    st = torch.cuda.current_stream(DEV).cuda_stream
    for args, msg in (((1, 0, 5, 5, 10, 10), 'low must be < high'),
                      ((1, 2, 0, 5, 10, 10), 'multiple of 4'),
                      ((1, 0, 0, 5, 0, 10), 'per_batch'),
                      ((1, 0, 0, 1 << 28, 10, 10), 'range must be')):
        with pytest.raises(_cabi.TGMNativeError, match=msg):
            _cabi.check(_cabi.lib.tgm_negatives_window(*args, out.data_ptr(), st))


def _graph(seed=3, N=500, E=4130, T=900, D=8):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, N // 2, E).astype(np.int32)
    dst = rng.integers(N // 2, N, E).astype(np.int32)
    t = np.sort(rng.integers(0, T, E)).astype(np.int64)
    x = rng.standard_normal((E, D)).astype(np.float32)
    data = DGData.from_raw(torch.from_numpy(t), torch.from_numpy(np.stack([src, dst], 1)),
                           torch.from_numpy(x))
    return DGraph(data, device=DEV), N


def _run(dg, N, nn, bs, window, stop_after=None):
    kw = {} if window is None else {'window_batches': window}
    hm = HookManager(keys=['k'])
    neg_hook = RandomNegativeEdgeSamplerHook(low=N // 2, high=N, **kw)
    hook = RecencyNeighborHook(num_nodes=N, num_nbrs=nn,
                               seed_nodes_keys=['edge_src', 'edge_dst', 'neg'],
                               seed_times_keys=['edge_time', 'edge_time', 'neg_time'], **kw)
    hm.register('k', neg_hook)
    hm.register('k', hook)
    out = []
    with hm.activate('k'):
        for b, batch in enumerate(DGDataLoader(dg, batch_size=bs, hook_manager=hm)):
            out.append(([v.clone() for v in (batch.neg, batch.neg_time)],
                        [[v.clone() for v in lst] for lst in (
                            batch.seed_nids, batch.seed_times, batch.nbr_nids,
                            batch.nbr_edge_time, batch.nbr_edge_x)],
                        {k_: v.clone() for k_, v in batch.seed_node_nbr_mask.items()}))
            if stop_after is not None and b + 1 == stop_after:
                break
        hm.reset_state()
    return out, hook, neg_hook


@pytest.mark.parametrize('nn', [[10], [5, 3]], ids=['1hop', '2hop'])
def test_default_hooks_serve_src_dst_neg_windows_identical_to_batch_by_batch(nn):
    dg, N = _graph()
    bs = 200  # 4130 edges: 20 full batches and a short one
    torch.manual_seed(7)
    ref, hook0, _ = _run(dg, N, nn, bs, 0)          # one randint + ring kernels per batch
    off_ref = _gen().get_offset()
    torch.manual_seed(7)
    got, hook, _ = _run(dg, N, nn, bs, None)        # default-constructed hooks
    assert _gen().get_offset() == off_ref
    assert len(got) == len(ref) == 21
    for b, (g, r) in enumerate(zip(got, ref)):
        for u, v in zip(g[0], r[0]):
            assert torch.equal(u, v), f'batch {b}: negatives differ'
        for lst_g, lst_r in zip(g[1], r[1]):
            for h, (u, v) in enumerate(zip(lst_g, lst_r)):
                assert u.dtype == v.dtype and torch.equal(u, v), f'batch {b} hop {h}'
        assert g[2].keys() == r[2].keys() == {'edge_src', 'edge_dst', 'neg'}
        for k_ in g[2]:
            assert torch.equal(g[2][k_], r[2][k_])
    torch.manual_seed(7)
    small, _, _ = _run(dg, N, nn, bs, 4)            # windows that end mid-stream
    for g, r in zip(small, ref):
        assert torch.equal(g[0][0], r[0][0])
        for lst_g, lst_r in zip(g[1], r[1]):
            for u, v in zip(lst_g, lst_r):
                assert torch.equal(u, v)


def test_windowed_run_stays_windowed_and_rewinds_the_generator_on_an_early_stop():
    dg, N = _graph()
    torch.manual_seed(11)
    off0 = _gen().get_offset()
    got, hook, neg_hook = _run(dg, N, [4], 200, None, stop_after=3)
    # three batches were served: per-batch draws would have advanced the generator by 3 * 4
    assert _gen().get_offset() == off0 + 12
    torch.manual_seed(11)
    want = [torch.randint(N // 2, N, (200,), dtype=torch.int32, device=DEV) for _ in range(3)]
    for g, w in zip(got, want):
        assert torch.equal(g[0][0], w)


def test_out_of_range_published_negatives_raise_like_the_reference():
    dg, N = _graph()
    hm = HookManager(keys=['k'])
    hm.register('k', RandomNegativeEdgeSamplerHook(low=0, high=N + 50))
    hm.register('k', RecencyNeighborHook(
        num_nodes=N, num_nbrs=[3], seed_nodes_keys=['edge_src', 'edge_dst', 'neg'],
        seed_times_keys=['edge_time', 'edge_time', 'neg_time']))
    torch.manual_seed(0)
    with hm.activate('k'), pytest.raises(ValueError, match='Seed nodes in neg must satisfy'):
        for _ in DGDataLoader(dg, batch_size=200, hook_manager=hm):
            pass


class _PublishFixtureNegatives:
    """The fixture's negatives (drawn by the reference run) handed out as a published window."""
    has_state = False
    requires = {'edge_src', 'edge_dst', 'edge_time'}
    produces = {'neg', 'neg_time'}

    def __init__(self, neg):
        self.neg, self.pub = neg, None

    def reset_state(self):
        pass

    def __call__(self, dg, batch):
        store, lo, hi = batch._slab[:3]
        if self.pub is None:
            self.pub = SeedWindow(store, 0, store.num_edges, self.neg, store._t.clone(), 0, 1 << 30)
        batch.neg, batch.neg_time = self.pub.nodes[lo:hi], self.pub.times[lo:hi]
        batch._seed_windows = {'neg': self.pub}
        return batch


@pytest.mark.parametrize('window', [None, 2])
@pytest.mark.parametrize('path', [p for p in golden_files() if Golden(p).neg is not None],
                         ids=[i for p, i in zip(golden_files(), golden_ids())
                              if Golden(p).neg is not None])
def test_published_negative_window_matches_reference_fixture(path, window):
    g = Golden(path)
    ei = torch.from_numpy(np.stack([g.src, g.dst], 1).astype(np.int32))
    dg = DGraph(DGData.from_raw(torch.from_numpy(g.t), ei,
                                None if g.x is None else torch.from_numpy(g.x)), device=DEV)
    hm = HookManager(keys=['g'])
    hm.register('g', _PublishFixtureNegatives(torch.from_numpy(g.neg.astype(np.int32)).to(DEV)))
    kw = {} if window is None else {'window_batches': window}
    hook = RecencyNeighborHook(num_nodes=g.N, num_nbrs=g.num_nbrs,
                               seed_nodes_keys=['edge_src', 'edge_dst', 'neg'],
                               seed_times_keys=['edge_time', 'edge_time', 'neg_time'],
                               directed=g.directed, **kw)
    hm.register('g', hook)
    with hm.activate('g'):
        for ep in range(g.epochs):
            for b, batch in enumerate(DGDataLoader(dg, batch_size=g.bs, hook_manager=hm)):
                assert isinstance(hook._win, dict) and hook._win['pub'] is not None
                for h in range(len(g.num_nbrs)):
                    got = tuple(v.cpu().numpy() for v in (
                        batch.seed_nids[h], batch.seed_times[h], batch.nbr_nids[h],
                        batch.nbr_edge_time[h], batch.nbr_edge_x[h]))
                    assert_hop_equal(got, g.expect(ep, b, h), f'ep{ep} batch{b} hop{h}')
                n = batch.edge_src.numel()
                assert batch.seed_node_nbr_mask['neg'].tolist() == list(range(2 * n, 3 * n))
            if ep + 1 < g.epochs:
                hm.reset_state()
