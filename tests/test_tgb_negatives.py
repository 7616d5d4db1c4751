"""TGB negative-sampler hooks (tgm_b200/hooks/tgb_negatives.py) against fixtures recorded from the
UNMODIFIED reference hook (tests/golden/make_golden_tgbneg.py; tgb_sampler.py:16-309) driven by the
same stand-in sampler (tests/_fake_tgb.py)."""
import glob
import os
import sys
import types

import numpy as np
import pytest
import torch

from tests._fake_tgb import FakeNegativeEdgeSampler
from tests._golden import GOLDEN_DIR
from tgm_b200 import DGData, DGDataLoader, DGraph, HookManager
from tgm_b200.hooks import (TGBNegativeEdgeSamplerHook, TGBTHGNegativeEdgeSamplerHook,
                            TGBTKGNegativeEdgeSamplerHook)

FIXTURES = sorted(glob.glob(os.path.join(GOLDEN_DIR, 'tgbneg_*.npz')))


def _hook(kind, N, **kw):
    s = FakeNegativeEdgeSampler(num_nodes=N)
    if kind == 'tgbl':
        return TGBNegativeEdgeSamplerHook('tgbl-fake', 'val', neg_sampler=s, **kw)
    if kind == 'thgl':
        return TGBTHGNegativeEdgeSamplerHook('thgl-fake', 'test', 0, N - 1,
                                             torch.arange(N, dtype=torch.int32) % 2,
                                             neg_sampler=s, **kw)
    return TGBTKGNegativeEdgeSamplerHook('tkgl-fake', 'val', 0, N - 1, neg_sampler=s, **kw)


def _graph(z, kind, device):
    kw = {}
    if int(z['with_type']):
        kw['edge_type'] = torch.from_numpy(z['et'])
        if kind == 'thgl':
            kw['node_type'] = torch.arange(int(z['N']), dtype=torch.int32) % 2
    data = DGData.from_raw(torch.from_numpy(z['t']),
                           torch.from_numpy(np.stack([z['src'], z['dst']], 1)), **kw)
    return DGraph(data, device=device)


def _run(z, kind, device, id=None):
    dg = _graph(z, kind, device)
    hm = HookManager(keys=['g'])
    hm.register('g', _hook(kind, int(z['N']), **({'id': id} if id else {})))
    with hm.activate('g'):
        return list(DGDataLoader(dg, batch_size=int(z['bs']), hook_manager=hm)), dg


def _run_host(z, kind, hook=None):
    """The hook alone over host batches (the package has no CPU data path: the loader needs a
    CUDA graph; the hook itself only reads the batch's tensors and dg.device / dg.num_nodes)."""
    hook = hook or _hook(kind, int(z['N']))
    dg = types.SimpleNamespace(device=torch.device('cpu'), num_nodes=int(z['N']))
    E, bs = len(z['src']), int(z['bs'])
    out = []
    for lo in range(0, E, bs):
        b = types.SimpleNamespace(edge_src=torch.from_numpy(z['src'][lo:lo + bs]),
                                  edge_dst=torch.from_numpy(z['dst'][lo:lo + bs]),
                                  edge_time=torch.from_numpy(z['t'][lo:lo + bs]),
                                  edge_type=torch.from_numpy(z['et'][lo:lo + bs]))
        out.append(hook(dg, b))
    return out


@pytest.mark.parametrize('path', FIXTURES, ids=lambda p: os.path.basename(p)[7:-4])
def test_tgb_hooks_match_the_reference_on_cpu(path):
    z = np.load(path)
    kind = os.path.basename(path)[7:-4]
    batches = _run_host(z, kind)
    assert len(batches) == int(z['nb'])
    for b, batch in enumerate(batches):
        assert batch.neg.dtype == torch.int32 and batch.neg_time.dtype == torch.int64
        assert np.array_equal(batch.neg.numpy(), z[f'b{b}_neg'])
        assert np.array_equal(batch.neg_time.numpy(), z[f'b{b}_neg_time'])  # same seeded CPU stream
        assert [x.numel() for x in batch.neg_batch_list] == z[f'b{b}_sizes'].tolist()
        assert all(x.dtype == torch.int32 for x in batch.neg_batch_list)
        assert np.array_equal(torch.cat(batch.neg_batch_list).numpy(), z[f'b{b}_flat'])


@pytest.mark.gpu
@pytest.mark.parametrize('path', FIXTURES, ids=lambda p: os.path.basename(p)[7:-4])
def test_tgb_hooks_match_the_reference_on_the_device(path):
    """ids and candidate lists equal the reference's; neg_time is the reference's own op on the
    device: torch.randint from a fresh generator seeded with 0 (tgb_sampler.py:120-129)."""
    z = np.load(path)
    kind = os.path.basename(path)[7:-4]
    batches, dg = _run(z, kind, 'cuda:0')
    lo = 0
    for b, batch in enumerate(batches):
        assert batch.neg.is_cuda and batch.neg.dtype == torch.int32
        assert np.array_equal(batch.neg.cpu().numpy(), z[f'b{b}_neg'])
        assert np.array_equal(torch.cat(batch.neg_batch_list).cpu().numpy(), z[f'b{b}_flat'])
        t = z['t'][lo:lo + batch.edge_src.numel()]
        gen = torch.Generator(device='cuda:0')
        gen.manual_seed(0)
        want = torch.randint(int(t.min()), int(t.max()) + 1, (batch.neg.numel(),), device='cuda:0',
                             generator=gen)
        assert torch.equal(batch.neg_time, want)
        lo += batch.edge_src.numel()


def test_tgb_hook_contract_and_errors():
    h = _hook('tgbl', 10)
    assert h.requires == {'edge_src', 'edge_dst', 'edge_time'}
    assert h.produces == {'neg', 'neg_batch_list', 'neg_time'}
    h = _hook('thgl', 10, id='foo')
    assert h.requires == {'edge_src', 'edge_dst', 'edge_time', 'edge_type'}
    assert h.produces == {'neg_foo', 'neg_batch_list_foo', 'neg_time_foo'} and 'foo' in repr(h)
    s = FakeNegativeEdgeSampler(num_nodes=10)
    with pytest.raises(ValueError, match='split_mode'):
        TGBNegativeEdgeSamplerHook('tgbl-fake', 'train', neg_sampler=s)
    with pytest.raises(ValueError, match='tgbl-xxx'):
        TGBNegativeEdgeSamplerHook('thgl-fake', 'val', neg_sampler=s)
    with pytest.raises(ValueError, match='positive'):
        TGBTKGNegativeEdgeSamplerHook('tkgl-fake', 'val', -1, 5, neg_sampler=s)
    with pytest.raises(ValueError, match='within node_type'):
        TGBTHGNegativeEdgeSamplerHook('thgl-fake', 'val', 0, 50, torch.zeros(3), neg_sampler=s)
    with pytest.raises(ValueError, match='must not be None'):
        TGBTHGNegativeEdgeSamplerHook('thgl-fake', 'val', 0, 5, None, neg_sampler=s)
    if 'tgb' not in sys.modules:
        with pytest.raises(ImportError, match='py-tgb'):
            TGBNegativeEdgeSamplerHook('tgbl-fake', 'val')


def test_tgb_hook_loads_the_evaluation_set_like_the_reference(monkeypatch):
    """Without an injected sampler the constructor imports tgb, builds the sampler and loads
    <PROJ_DIR>datasets/<name>/<name>_<split>_ns[_vN].pkl (tgb_sampler.py:48-77)."""
    mods = {n: types.ModuleType(n) for n in
            ['tgb', 'tgb.utils', 'tgb.utils.info', 'tgb.linkproppred',
             'tgb.linkproppred.negative_sampler']}
    mods['tgb.utils.info'].DATA_VERSION_DICT = {'tgbl-fake': 2}
    mods['tgb.utils.info'].PROJ_DIR = '/nonexistent/'
    mods['tgb.linkproppred.negative_sampler'].NegativeEdgeSampler = FakeNegativeEdgeSampler
    for n, m in mods.items():
        monkeypatch.setitem(sys.modules, n, m)
    h = TGBNegativeEdgeSamplerHook('tgbl-fake', 'val')
    assert h.neg_sampler.loaded == ('/nonexistent/datasets/tgbl_fake/tgbl-fake_val_ns_v2.pkl', 'val')
    h = TGBNegativeEdgeSamplerHook('tgbl-other', 'test')
    assert h.neg_sampler.loaded == ('/nonexistent/datasets/tgbl_other/tgbl-other_test_ns.pkl', 'test')


def test_tgb_hook_reports_sampler_failures_like_the_reference():
    class Broken(FakeNegativeEdgeSampler):
        def query_batch(self, *a, **k):
            raise ValueError('edge not in eval set')

    z = np.load(FIXTURES[0])
    with pytest.raises(ValueError, match='TGBL Negative sampling failed'):
        _run_host(z, 'tgbl', TGBNegativeEdgeSamplerHook('tgbl-fake', 'val', neg_sampler=Broken()))
    # an empty batch asks the sampler nothing (tgb_sampler.py:93-100)
    dg = types.SimpleNamespace(device=torch.device('cpu'), num_nodes=5)
    e32, e64 = torch.empty(0, dtype=torch.int32), torch.empty(0, dtype=torch.int64)
    b = TGBNegativeEdgeSamplerHook('tgbl-fake', 'val', neg_sampler=Broken())(
        dg, types.SimpleNamespace(edge_src=e32, edge_dst=e32, edge_time=e64))
    assert b.neg.numel() == 0 and b.neg.dtype == torch.int32 and b.neg_batch_list == []
    assert b.neg_time.numel() == 0 and b.neg_time.dtype == torch.int64
