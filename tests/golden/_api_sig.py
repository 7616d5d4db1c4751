"""Signature description shared by make_golden_api.py (reference side) and the test (drop-in side)."""
from __future__ import annotations

import inspect


def resolve(root, dotted):
    import importlib
    obj, parts = root, dotted.split('.')
    for i, p in enumerate(parts):
        if not hasattr(obj, p):
            obj = importlib.import_module(root.__name__ + '.' + '.'.join(parts[:i + 1]))
        else:
            obj = getattr(obj, p)
    return obj


def describe(fn):
    out = []
    for name, p in inspect.signature(fn).parameters.items():
        d = p.default
        if d is inspect.Parameter.empty:
            default = ['required']
        elif isinstance(d, (int, float, str, bool, type(None))):
            default = ['value', d]
        else:
            default = ['object', getattr(d, '__name__', type(d).__name__)]
        out.append([name, p.kind.name, default])
    return out
