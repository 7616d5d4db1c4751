"""pytest plugin (test infrastructure): makes `import tgm...` resolve to `tgm_b200...`, so the
reference's OWN unit tests can be run against the drop-in's host-side logic
(tests/test_reference_suite_differential.py).  Loaded with `-p tests._reference_alias_plugin`."""
import importlib
import os
import pkgutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import tgm_b200  # noqa: E402


def _alias(pkg, as_name: str) -> None:
    sys.modules[as_name] = pkg
    for m in pkgutil.iter_modules(getattr(pkg, '__path__', [])):
        if m.name in ('csrc', 'build'):
            continue
        _alias(importlib.import_module(f'{pkg.__name__}.{m.name}'), f'{as_name}.{m.name}')


_alias(tgm_b200, 'tgm')
# import paths the package registers without a file behind them (tgm_b200/nn/__init__.py)
for _name, _m in list(sys.modules.items()):
    if _name.startswith('tgm_b200.') and _m is not None:
        sys.modules.setdefault('tgm' + _name[len('tgm_b200'):], _m)


# Names the reference's test modules import that are OUT OF SCOPE here (splits,
# TGB loaders/recipes, historical / pre-generated negatives): placeholders that refuse to be used,
# so that the in-scope tests of those modules can still be collected and run.
import types  # noqa: E402

import tgm_b200.data as _data  # noqa: E402
import tgm_b200.hooks as _hooks  # noqa: E402


class _OutOfScope:
    def __init__(self, *a, **k):
        raise NotImplementedError('out of scope for tgm_b200')


for _mod, _names in ((_data, ['TemporalRatioSplit', 'TemporalSplit', 'TGBSplit', 'SplitStrategy']),
                     (_hooks, ['TGBNegativeEdgeSamplerHook', 'HistoricalNegativeEdgeSamplerHook',
                               'RecipeRegistry'])):
    for _n in _names:
        if not hasattr(_mod, _n):
            setattr(_mod, _n, type(_n, (_OutOfScope,), {}))
_split = types.ModuleType('tgm.data.split')
for _n in ['TemporalRatioSplit', 'TemporalSplit', 'TGBSplit', 'SplitStrategy']:
    setattr(_split, _n, getattr(_data, _n))
sys.modules['tgm.data.split'] = _split

import tgm_b200.core.timedelta as _td  # noqa: E402

for _n in ('TGB_TIME_DELTAS', 'TGB_SEQ_TIME_DELTAS'):  # per-dataset tables of the TGB loaders
    if not hasattr(_td, _n):
        setattr(_td, _n, {})
