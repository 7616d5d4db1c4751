"""Import shim that makes the *unmodified* reference (tgm-team/tgm under /root/reference)
importable in the build container, where torch_geometric / tgb are not installed.

Only used by tests/golden/make_golden.py (fixture generation) and by the optional
`-m "not gpu"` cross-checks that skip themselves when /root/reference is absent
(it does not exist on the GPU box).  Nothing in the product path imports this.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get('TGM_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'tgm'))


def _stub_torch_geometric() -> None:
    import torch

    if 'torch_geometric' in sys.modules:
        return

    class _Dummy(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    def _noop(*a, **k):
        return None

    def _scatter(src, index, dim=0, dim_size=None, reduce='sum'):
        # restatement of torch_geometric.utils.scatter for the 1-D/2-D dim=0 cases TGN uses
        red = {'sum': 'sum', 'add': 'sum', 'mean': 'mean', 'max': 'amax', 'min': 'amin'}[reduce]
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() else 0
        out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
        idx = index.long()
        if src.dim() > 1:
            idx = idx.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
        return out.scatter_reduce_(0, idx, src, reduce=red, include_self=False)

    names = ['torch_geometric', 'torch_geometric.nn', 'torch_geometric.nn.inits',
             'torch_geometric.nn.models', 'torch_geometric.nn.models.tgn',
             'torch_geometric.utils']
    mods = {n: types.ModuleType(n) for n in names}
    for cls in ['GCNConv', 'Linear', 'AntiSymmetricConv', 'TransformerConv', 'ChebConv']:
        setattr(mods['torch_geometric.nn'], cls, type(cls, (_Dummy,), {}))
    def _zeros(value):
        # torch_geometric.nn.inits.zeros: in-place fill with 0 (TGNMemory.reset_state relies on it,
        # tgm/nn/encoder/tgn.py:149-152)
        if isinstance(value, torch.Tensor):
            value.data.fill_(0)

    for fn in ['ones', 'glorot']:
        setattr(mods['torch_geometric.nn.inits'], fn, _noop)
    mods['torch_geometric.nn.inits'].zeros = _zeros
    mods['torch_geometric.nn.models.tgn'].TimeEncoder = type('TimeEncoder', (_Dummy,), {})
    mods['torch_geometric.utils'].scatter = _scatter
    mods['torch_geometric'].nn = mods['torch_geometric.nn']
    mods['torch_geometric'].utils = mods['torch_geometric.utils']
    mods['torch_geometric.nn'].inits = mods['torch_geometric.nn.inits']
    mods['torch_geometric.nn'].models = mods['torch_geometric.nn.models']
    mods['torch_geometric.nn.models'].tgn = mods['torch_geometric.nn.models.tgn']
    sys.modules.update(mods)


def import_reference():
    """Return the reference `tgm` package (raises ImportError when not present)."""
    if not reference_available():
        raise ImportError(f'reference tree not found under {REFERENCE_ROOT}')
    _stub_torch_geometric()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import tgm  # noqa: F401

    return tgm
