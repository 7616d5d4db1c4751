"""Fixtures for the TGB negative-sampler hooks, from the UNMODIFIED reference hook:

    python tests/golden/make_golden_tgbneg.py      # build container only

tgm/hooks/negatives/tgb_sampler.py:16-309 imports the third-party `tgb` package (py-tgb, absent
here and on the GPU box); a stand-in package whose samplers are tests/_fake_tgb.py is installed in
sys.modules, so the hook's own code (`__init__`, `_query_batch`, `__call__`) runs unchanged on CPU.
Writes tests/golden/tgbneg_{tgbl,thgl,tkgl}.npz: per batch neg, neg_time and the candidate lists.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _ref_shim import import_reference  # noqa: E402
from tests._fake_tgb import FakeNegativeEdgeSampler  # noqa: E402

N = 50
for name in ['tgb', 'tgb.utils', 'tgb.utils.info', 'tgb.linkproppred',
             'tgb.linkproppred.negative_sampler', 'tgb.linkproppred.thg_negative_sampler',
             'tgb.linkproppred.tkg_negative_sampler']:
    sys.modules[name] = types.ModuleType(name)
sys.modules['tgb.utils.info'].DATA_VERSION_DICT = {'tgbl-fake': 2}
sys.modules['tgb.utils.info'].PROJ_DIR = '/nonexistent/'
mk = lambda: type('S', (FakeNegativeEdgeSampler,), {  # noqa: E731
    '__init__': lambda self, **kw: FakeNegativeEdgeSampler.__init__(self, num_nodes=N, **kw)})
sys.modules['tgb.linkproppred.negative_sampler'].NegativeEdgeSampler = mk()
sys.modules['tgb.linkproppred.thg_negative_sampler'].THGNegativeEdgeSampler = mk()
sys.modules['tgb.linkproppred.tkg_negative_sampler'].TKGNegativeEdgeSampler = mk()

import_reference()
from tgm import DGraph  # noqa: E402
from tgm.data import DGData, DGDataLoader  # noqa: E402
from tgm.hooks import (HookManager, TGBNegativeEdgeSamplerHook,  # noqa: E402
                       TGBTHGNegativeEdgeSamplerHook, TGBTKGNegativeEdgeSamplerHook)


def save(kind, hook, with_type, bs=7, E=40):
    rng = np.random.default_rng(hash(kind) % 1000)
    src = rng.integers(0, N, E).astype(np.int32)
    dst = rng.integers(0, N, E).astype(np.int32)
    t = np.sort(rng.integers(0, 500, E)).astype(np.int64)
    et = rng.integers(0, 3, E).astype(np.int32)
    kw = {}
    if with_type:
        kw['edge_type'] = torch.from_numpy(et)
        if kind == 'thgl':
            kw['node_type'] = torch.arange(N, dtype=torch.int32) % 2
    data = DGData.from_raw(torch.from_numpy(t), torch.from_numpy(np.stack([src, dst], 1)), **kw)
    dg = DGraph(data)
    hm = HookManager(keys=['g'])
    hm.register('g', hook)
    out, nb = {}, 0
    with hm.activate('g'):
        for batch in DGDataLoader(dg, batch_size=bs, hook_manager=hm):
            out[f'b{nb}_neg'] = batch.neg.numpy()
            out[f'b{nb}_neg_time'] = batch.neg_time.numpy()
            out[f'b{nb}_sizes'] = np.array([x.numel() for x in batch.neg_batch_list], np.int64)
            out[f'b{nb}_flat'] = torch.cat(batch.neg_batch_list).numpy()
            assert batch.neg.dtype == torch.int32 and batch.neg_time.dtype == torch.int64
            nb += 1
    np.savez_compressed(os.path.join(HERE, f'tgbneg_{kind}.npz'), src=src, dst=dst, t=t, et=et,
                        N=np.int64(N), bs=np.int64(bs), nb=np.int64(nb),
                        with_type=np.int64(with_type), **out)
    print(kind, nb, 'batches; sampler loaded', hook.neg_sampler.loaded)


save('tgbl', TGBNegativeEdgeSamplerHook('tgbl-fake', 'val'), False)
save('thgl', TGBTHGNegativeEdgeSamplerHook('thgl-fake', 'test', 0, N - 1,
                                           torch.arange(N, dtype=torch.int32) % 2), True)
save('tkgl', TGBTKGNegativeEdgeSamplerHook('tkgl-fake', 'val', 0, N - 1), True)
