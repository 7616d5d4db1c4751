"""Host-side mirror of the reference interface: DGData validation, view algebra, loader batch
boundaries, hook protocol and dependency ordering.  CPU only (metadata-only stores)."""
import json
import os
import warnings

import numpy as np
import pytest
import torch

from tests._golden import GOLDEN_DIR

from tgm_b200 import (DGBatch, DGData, DGDataLoader, DGraph, HookManager, RecencyNeighborHook,
                      RandomNegativeEdgeSamplerHook, DeduplicationHook, TimeDeltaDG, _cabi)
from tgm_b200.core.storage import DGSliceTracker
from tgm_b200.exceptions import (BadHookProtocolError, EmptyGraphError, EventOrderedConversionError,
                                 InvalidNodeIDError, UnresolvableHookDependenciesError)
from tgm_b200.hooks.base import DGHook, StatelessHook


def _data(n=10, D=2, time_delta='r'):
    ei = torch.stack([torch.arange(n) % 4, (torch.arange(n) + 1) % 5], 1).int()
    t = torch.tensor([1, 1, 2, 3, 3, 3, 5, 8, 8, 9][:n])
    return DGData.from_raw(t, ei, torch.arange(n * D).view(n, D).float(), time_delta=time_delta)


# --- DGData (dg_data.py:86-394) -----------------------------------------------------------
def test_dgdata_casts_and_sorts():
    ei = torch.tensor([[0, 1], [2, 3], [1, 2]])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        d = DGData.from_raw(torch.tensor([5, 1, 3], dtype=torch.int32), ei,
                            torch.tensor([[5.], [1.], [3.]], dtype=torch.float64))
    assert d.time.dtype == torch.int64 and d.time.tolist() == [1, 3, 5]
    assert d.edge_index.dtype == torch.int32 and d.edge_index.tolist() == [[2, 3], [1, 2], [0, 1]]
    assert d.edge_x.dtype == torch.float32 and d.edge_x[:, 0].tolist() == [1., 3., 5.]
    assert d.num_nodes == 4 and d.edge_mask.tolist() == [0, 1, 2]


def test_dgdata_rejects_bad_input():
    ei = torch.tensor([[0, 1]], dtype=torch.int32)
    with pytest.raises(InvalidNodeIDError):
        DGData.from_raw(torch.tensor([1]), torch.tensor([[0, -1]], dtype=torch.int32))
    with pytest.raises(ValueError):
        DGData.from_raw(torch.tensor([-1]), ei)
    with pytest.raises(ValueError):
        DGData.from_raw(torch.tensor([2 ** 31 - 1]), ei)
    with pytest.raises(TypeError):
        DGData.from_raw(torch.tensor([1.5]), ei)
    with pytest.raises(EmptyGraphError):
        DGData.from_raw(torch.empty(0, dtype=torch.int64), torch.empty(0, 2, dtype=torch.int32))
    with pytest.raises(ValueError):
        DGData.from_raw(torch.tensor([1]), ei, torch.zeros(2, 3))
    with pytest.raises(TypeError):
        DGraph('not data')


def test_node_events_share_the_timeline():
    ei = torch.tensor([[0, 1], [1, 2]], dtype=torch.int32)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        d = DGData.from_raw(torch.tensor([2, 6]), ei, node_x_time=torch.tensor([1, 4]),
                            node_x_nids=torch.tensor([3, 0], dtype=torch.int32),
                            node_x=torch.ones(2, 3))
    assert d.time.tolist() == [1, 2, 4, 6]
    assert d.edge_mask.tolist() == [1, 3] and d.node_x_mask.tolist() == [0, 2]
    dg = DGraph(d)
    assert dg.num_events == 4 and dg.num_edge_events == 2 and dg.num_node_events == 2
    assert dg.num_nodes == 4
    sub = dg.slice_time(0, 3)  # events at t in [0, 3)
    assert sub.num_events == 2 and sub.num_edge_events == 1
    nx = sub.node_x
    assert nx is not None and nx._indices().tolist() == [[1], [3]]


# --- view algebra (graph.py:110-152) ------------------------------------------------------
def test_slice_time_is_end_exclusive():
    dg = DGraph(_data())
    assert (dg.start_time, dg.end_time, dg.num_events, dg.num_timestamps) == (1, 9, 10, 6)
    s = dg.slice_time(3, 8)
    assert s._slice == DGSliceTracker(start_time=3, end_time=7)
    assert (s.num_events, s.start_time, s.end_time) == (4, 3, 7)
    assert s._storage.edge_range(s._slice) == (3, 7)
    assert dg.slice_time(8, 9).num_events == 2
    with pytest.raises(ValueError):
        dg.slice_time(5, 4)


def test_slice_events_and_composition():
    dg = DGraph(_data())
    s = dg.slice_events(2, 7)
    assert s.num_events == 5 and s._storage.edge_range(s._slice) == (2, 7)
    s2 = s.slice_events(0, 5)  # intersection, not re-basing (graph.py:124-126)
    assert s2._storage.edge_range(s2._slice) == (2, 5)
    s3 = s.slice_time(3, 4)
    assert s3._storage.edge_range(s3._slice) == (3, 6)
    assert dg.slice_events(4, 4).num_events == 0
    assert dg.slice_events(4, 4).start_time is None
    with pytest.raises(ValueError):
        dg.slice_events(3, 2)


def test_cpu_view_refuses_edge_data():
    dg = DGraph(_data())
    with pytest.raises(_cabi.TGMNativeError):
        dg.materialize()


# --- loader (loader.py:101-170) -----------------------------------------------------------
def test_loader_batch_counts():
    dg = DGraph(_data())
    assert len(DGDataLoader(dg, batch_size=3)) == 4
    assert len(DGDataLoader(dg, batch_size=3, drop_last=True)) == 3
    assert len(DGDataLoader(dg, batch_size=10)) == 1
    with pytest.raises(ValueError):
        DGDataLoader(dg, batch_size=0)
    with pytest.raises(ValueError):
        DGDataLoader(dg, on_empty='nope')
    with pytest.raises(EventOrderedConversionError):
        DGDataLoader(dg, batch_size=1, batch_unit='s')
    dgt = DGraph(_data(time_delta='s'))
    assert len(DGDataLoader(dgt, batch_size=2, batch_unit='s')) == 5  # t in [1, 10) by 2


def test_timedelta():
    assert TimeDeltaDG('r').is_event_ordered and TimeDeltaDG('s').is_time_ordered
    assert TimeDeltaDG('m').convert('s') == 60 and TimeDeltaDG('s').convert('m') == 1 / 60
    assert TimeDeltaDG('h').is_coarser_than('m') and not TimeDeltaDG('s').is_coarser_than('s')
    for bad in [('r', 2), ('s', 0), ('parsec', 1)]:
        with pytest.raises(ValueError):
            TimeDeltaDG(*bad)
    with pytest.raises(EventOrderedConversionError):
        TimeDeltaDG('r').convert('s')


# --- hooks (hooks/base.py, hook_manager.py) -----------------------------------------------
def _nbr_hook(**kw):
    args = dict(num_nodes=5, num_nbrs=[2], seed_nodes_keys=['edge_src', 'edge_dst'],
                seed_times_keys=['edge_time', 'edge_time'])
    args.update(kw)
    return RecencyNeighborHook(**args)


def test_recency_hook_contract():
    """test_recency_nbr_hook.py:52-96,195-247 of the reference."""
    h = _nbr_hook()
    assert isinstance(h, DGHook) and h.has_state
    assert h.requires == {'edge_src', 'edge_dst', 'edge_time'}
    assert h.produces == {'seed_nids', 'seed_times', 'nbr_nids', 'nbr_edge_time', 'nbr_edge_x',
                          'seed_node_nbr_mask'}
    hid = _nbr_hook(id='foo', seed_nodes_keys=['edge_src', 'neg'])
    assert hid.requires == {'edge_src', 'edge_dst', 'edge_time', 'neg'}
    assert hid.produces == {f'{p}_foo' for p in h.produces} and 'foo' in repr(hid)
    assert h.num_nbrs == [2]
    for bad in [dict(num_nbrs=[]), dict(num_nbrs=[0]), dict(num_nbrs=[1.5]), dict(num_nbrs=[-1]),
                dict(seed_times_keys=['edge_time'])]:
        with pytest.raises(ValueError):
            _nbr_hook(**bad)


def test_negative_hook_args():
    with pytest.raises(ValueError):
        RandomNegativeEdgeSamplerHook(low=0, high=3, neg_ratio=0)
    with pytest.raises(ValueError):
        RandomNegativeEdgeSamplerHook(low=3, high=3)
    h = RandomNegativeEdgeSamplerHook(low=0, high=3, id='x')
    assert h.produces == {'neg_x', 'neg_time_x'} and not h.has_state


class _Produces(StatelessHook):
    def __init__(self, requires, produces):
        self._init_hook()
        self._requires, self._produces = set(requires), set(produces)

    def __call__(self, dg, batch):
        batch.trace = getattr(batch, 'trace', []) + [sorted(self._produces)[0]]
        return batch


def test_hook_manager_orders_by_dependencies():
    hm = HookManager(keys=['train', 'val'])
    a, b, c = _Produces([], ['x']), _Produces(['x'], ['y']), _Produces(['y', 'x'], ['z'])
    hm.register('train', c)
    hm.register('train', b)
    hm.register_shared(a)
    with hm.activate('train'):
        out = hm.execute_active_hooks(None, DGBatch(None, None, None))
    assert out.trace == ['x', 'y', 'z']
    with hm.activate('val'):
        assert [type(h) for h in hm.active_hooks()] == [_Produces] and hm.active_hooks()[0] is a
    assert hm._active_key is None


def test_negatives_run_before_neighbours():
    """hook_manager.py:420-430: no data dependency links them, the order is forced."""
    hm = HookManager(keys=['k'])
    hm.register('k', _nbr_hook())
    hm.register('k', RandomNegativeEdgeSamplerHook(low=0, high=4))
    hm.register('k', DeduplicationHook(seed_nodes_keys=['neg', 'nbr_nids']))
    hm.set_active_hooks('k')
    names = [type(h).__name__ for h in hm.active_hooks()]
    assert names == ['RandomNegativeEdgeSamplerHook', 'RecencyNeighborHook', 'DeduplicationHook']


def test_hook_manager_errors():
    with pytest.raises(ValueError):
        HookManager(keys=[])
    hm = HookManager(keys=['k'])
    with pytest.raises(BadHookProtocolError):
        hm.register('k', object())
    with pytest.raises(KeyError):
        hm.register('nope', _nbr_hook())
    with pytest.raises(RuntimeError):
        hm.execute_active_hooks(None, None)
    hm.register('k', _Produces(['missing'], ['x']))
    with pytest.raises(UnresolvableHookDependenciesError):
        hm.resolve_hooks('k')
    hm2 = HookManager(keys=['k'])
    hm2.register('k', _Produces(['b'], ['a']))
    hm2.register('k', _Produces(['a'], ['b']))
    with pytest.raises(UnresolvableHookDependenciesError):
        hm2.resolve_hooks()
    with hm2.activate('k'):
        with pytest.raises(RuntimeError):
            hm2.register('k', _nbr_hook())


def test_validate_requirement_suggestions():
    class Enc:
        requires = {'nbr_nids', 'nbr_nidz', 'totally_unknown'}

        def __call__(self, batch):
            return batch

    hm = HookManager(keys=['k'])
    hm.register('k', RandomNegativeEdgeSamplerHook(low=0, high=4))
    with pytest.raises(UnresolvableHookDependenciesError) as e:
        hm.validate_requirement(Enc(), 'k')
    msg = str(e.value)
    assert "register 'RecencyNeighborHook'" in msg and "Do you mean 'nbr_nids'" in msg
    assert "'totally_unknown': Can not find" in msg
    hm.register('k', _nbr_hook())
    Enc.requires = {'nbr_nids', 'neg', 'edge_src'}
    hm.validate_requirement(Enc())


# --- view algebra + loader batch plan vs fixtures from the unmodified reference ---------------
def _loader_cases():
    z = np.load(os.path.join(GOLDEN_DIR, 'loader_plans.npz'))
    return z, sorted({k.split('/')[0] for k in z.files})


@pytest.mark.parametrize('name', _loader_cases()[1])
def test_view_metadata_and_loader_plan_match_reference_fixture(name):
    """tests/golden/make_golden_loader.py ran the reference's DGraph views and DGDataLoader
    (on_empty=None, so empty batches are recorded too) on CPU; the host side of the B200 store must
    put the same events into the same batches: edge events as the slab [lo, hi) of the sorted
    stream, node events and node labels by id and time.  (Materialising the slabs needs the GPU:
    tests/test_gpu_parity.py.)  Cases that are not `exact` have timestamp ties between events the
    reference re-orders with an unstable argsort (dg_data.py:351-358): they use time-window batches
    and are compared per batch as sorted sets; this store keeps the input order among ties."""
    z, _ = _loader_cases()
    g = lambda k: z[f'{name}/{k}']
    meta = json.loads(bytes(g('meta')).decode())
    spec = json.loads(bytes(g('spec')).decode())
    kw = dict(edge_time=torch.from_numpy(g('raw_t')),
              edge_index=torch.from_numpy(np.stack([g('raw_src'), g('raw_dst')], 1)),
              edge_x=torch.zeros(len(g('raw_t')), 2))
    for p in ('nx', 'ny'):
        if f'{name}/raw_{p}_t' in z.files:
            full = 'node_x' if p == 'nx' else 'node_y'
            kw.update({f'{full}_time': torch.from_numpy(g(f'raw_{p}_t')),
                       f'{full}_nids': torch.from_numpy(g(f'raw_{p}_id')),
                       full: torch.from_numpy(g(f'raw_{p}'))})
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        dg = DGraph(DGData.from_raw(time_delta=spec['time_delta'], **kw))
    for op, a, b in spec['ops']:
        dg = getattr(dg, op)(a, b)
    for key in ('start_time', 'end_time', 'num_events', 'num_edge_events', 'num_node_events',
                'num_node_labels', 'num_timestamps', 'num_nodes'):
        assert getattr(dg, key) == meta[key], key
    assert sorted(int(v) for v in dg._storage.get_nodes(dg._slice)) == meta['nodes']
    loader = DGDataLoader(dg, on_empty=None, **spec['loader'])
    assert len(loader) == meta['len']
    e_off, nx_off, ny_off = g('e_off'), g('nx_off'), g('ny_off')
    for i, start in enumerate(loader._starts):
        sub = loader._slice_op(start, start + loader._batch_size)
        lo, hi = sub._storage.edge_range(sub._slice)
        want = g('eids')[e_off[i]:e_off[i + 1]]
        order = (lambda v: v) if spec['exact'] else np.sort
        assert hi - lo == len(want) and np.array_equal(np.arange(lo, hi), order(want)), (i, lo, hi)
        for off, ids_k, t_k, getter in ((nx_off, 'nx_ids', 'nx_t', sub._storage.get_node_events),
                                        (ny_off, 'ny_ids', 'ny_t', sub._storage.get_node_labels)):
            nid, tt = (np.asarray(v, np.int64) for v in getter(sub._slice))
            w_id, w_t = g(ids_k)[off[i]:off[i + 1]], g(t_k)[off[i]:off[i + 1]]
            if not spec['exact']:  # order among equal times is implementation-defined upstream
                nid, tt = nid[np.lexsort((nid, tt))], np.sort(tt)
                w_id, w_t = w_id[np.lexsort((w_id, w_t))], np.sort(w_t)
            assert np.array_equal(nid, w_id) and np.array_equal(tt, w_t), (i, ids_k)


# --- public signatures vs a snapshot of the reference's -----------------------------------------
def test_public_signatures_accept_every_reference_argument():
    """tests/golden/api_signatures.json (make_golden_api.py) holds parameter names, order, kinds and
    defaults of the reference's constructors/methods on the hot path.  The drop-in must accept each
    of them at the same position with the same default; extra optional parameters may follow
    (e.g. RecencyNeighborHook(window_batches=...))."""
    import tgm_b200
    from tests.golden._api_sig import describe, resolve
    snap = json.load(open(os.path.join(GOLDEN_DIR, 'api_signatures.json')))
    assert len(snap) >= 35
    for dotted, want in snap.items():
        got = describe(resolve(tgm_b200, dotted))
        assert len(got) >= len(want), dotted
        for g, w in zip(got, want):
            assert g[:2] == w[:2], (dotted, g, w)  # same name, same kind, same position
            if w[2] == ['required']:  # may have become optional here (default None): still accepted
                assert g[2] == ['required'] or g[2] == ['value', None], (dotted, g, w)
            elif w[2][0] == 'object':  # a class / factory default upstream: any default will do
                assert g[2] != ['required'], (dotted, g, w)
            else:
                assert g[2] == w[2], (dotted, g, w)
        for extra in got[len(want):]:
            assert extra[2] != ['required'] or extra[1].startswith('VAR_'), (dotted, extra)


def test_reference_rng_picks_replay_the_reference_draws():
    """tgm_b200.sampler.reference_rng_picks (host half of NeighborSamplerHook(reference_rng=True))
    consumes CPython's global generator exactly like get_nbrs' random.sample(candidates, k)
    (array_backend.py:147-153): same picks as sampling the candidate lists themselves."""
    import random

    from tgm_b200.sampler import reference_rng_picks
    rng = np.random.default_rng(3)
    counts = rng.integers(0, 40, 200).tolist()
    for k in (1, 5, 21):  # 21 < count crosses CPython's set/pool switch (setsize 21 for k <= 5...)
        random.seed(99)
        got = reference_rng_picks(counts, k)
        random.seed(99)
        for c, row in zip(counts, got):
            cand = [('cand', i) for i in range(c)]
            want = random.sample(cand, k) if c > k else cand
            assert [p for p in row if p >= 0] == [i for _, i in want]
            assert len(row) == k and all(p == -1 for p in row[len(want):])


def test_block_views_equal_tensor_split():
    """`block_views` (core/storage.py) is Tensor.split(n) with cheaper views where the rows divide
    evenly: the same views (same memory, same shapes) in every case, one view per batch."""
    from tgm_b200.core.storage import block_views
    for rows, n, tail in [(12, 4, (2,)), (12, 5, (2,)), (4, 4, ()), (3, 4, (7, 2)), (4096, 200, (3,))]:
        x = torch.arange(rows * max(1, int(np.prod(tail)))).reshape(rows, *tail)
        got, want = block_views(x, n), x.split(n)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert g.shape == w.shape and g.data_ptr() == w.data_ptr() and torch.equal(g, w)


def test_reference_plugin_needs_the_reference_package():
    """tgm_b200.reference_plugin imports without the reference installed; `install()` is what
    needs `tgm` (here absent from sys.path unless the differential tests put it there)."""
    import importlib
    import sys
    mod = importlib.import_module('tgm_b200.reference_plugin')
    assert mod.BACKEND_NAME == 'B200Storage'
    if 'tgm' not in sys.modules:
        with pytest.raises(ImportError):
            mod.make_backend(None)
