"""Hook registry: `@hook` makes a hook class discoverable, `list_hooks()` returns the classes
(tgm/hooks/registry.py:8-22).  HookManager.validate_requirement searches it for suggestions."""
from __future__ import annotations

from typing import List

_HOOK_REGISTRY: List[type] = []


def hook(cls: type) -> type:
    """Class decorator registering a hook class."""
    _HOOK_REGISTRY.append(cls)  # as upstream: registering a class twice lists it twice
    return cls


def list_hooks() -> List[type]:
    return _HOOK_REGISTRY
