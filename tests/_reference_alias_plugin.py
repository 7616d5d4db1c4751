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
