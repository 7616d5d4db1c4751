"""HookManager: keyed + shared hooks, dependency-ordered execution.

Behavioural mirror of tgm/hooks/hook_manager.py:38-462 (host orchestration, kept as is by
SURVEY.md section 2): lazy Kahn topological sort on requires/produces per key (:389-462), the
hard-wired "negatives before neighbours" edge (:420-430), `activate` context (:214-226),
`execute_active_hooks` (:139-168), `reset_state` (:170-189), `validate_requirement` (:228-371).
"""
from __future__ import annotations

import difflib
from collections import deque
from contextlib import contextmanager
from typing import Any, Dict, Iterator, List, Optional, Set

from tgm_b200.exceptions import (BadEncoderProtocolError, BadHookProtocolError,
                                 UnresolvableHookDependenciesError)
from tgm_b200.hooks.base import DGHook
from tgm_b200.hooks.registry import hook as register_hook_class  # noqa: F401  (older name)
from tgm_b200.hooks.registry import list_hooks  # module-level name so tests can patch it

# attributes every materialised batch carries without any hook (:23-35)
CORE_ATTRIBUTE: Set[str] = {'edge_src', 'edge_dst', 'edge_time', 'edge_type', 'node_x_time',
                            'node_x_nids', 'node_y_time', 'node_y_nids', 'node_type'}


class HookManager:
    def __init__(self, keys: List[str]) -> None:
        if not len(keys):
            raise ValueError('HookManager keys list must be non-empty')
        self._registered_key = keys
        self._key_to_hooks: Dict[str, List[DGHook]] = {k: [] for k in keys}
        self._shared_hooks: List[DGHook] = []
        self._dirty: Dict[str, bool] = {k: False for k in keys}
        self._active_key: Optional[str] = None

    @property
    def keys(self) -> List[str]:
        return self._registered_key

    def __str__(self) -> str:
        def line(h) -> str:
            return f'    - {h!r} (requires={h.requires}, produces={h.produces})'
        out = ['HookManager:', '  Shared hooks:', *map(line, self._shared_hooks),
               f'  Active key: {self._active_key}', '  Keyed hooks:']
        for key, hooks in self._key_to_hooks.items():
            out += [f'    {key}:', *map(line, hooks)]
        return '\n'.join(out)

    # -- registration -----------------------------------------------------------------------
    def register_shared(self, hook: DGHook) -> None:
        self._check_hook(hook)
        self._check_inactive()
        self._shared_hooks.append(hook)
        for k in self._dirty:
            self._dirty[k] = True

    def register(self, key: str, hook: DGHook) -> None:
        self._check_key(key)
        self._check_hook(hook)
        self._check_inactive()
        self._key_to_hooks[key].append(hook)
        self._dirty[key] = True

    def set_active_hooks(self, key: str) -> None:
        self._check_key(key)
        self._active_key = key

    @contextmanager
    def activate(self, key: str) -> Iterator[None]:
        previous = self._active_key
        self.set_active_hooks(key)
        try:
            yield
        finally:
            self._active_key = previous

    # -- execution --------------------------------------------------------------------------
    def active_hooks(self) -> List[DGHook]:
        """Resolved execution order for the active key."""
        if self._active_key is None:
            raise RuntimeError('No active key set. Use activate() context manager.')
        key = self._active_key
        if self._dirty[key]:
            self.resolve_hooks(key)
        return self._key_to_hooks[key]

    def execute_active_hooks(self, dg, batch):
        for hook in self.active_hooks():
            batch = hook(dg, batch)
        return batch

    def reset_state(self, key: Optional[str] = None) -> None:
        if key is not None:
            self._check_key(key)
        for h in self._shared_hooks:
            h.reset_state()
        for k in ([key] if key is not None else list(self._key_to_hooks)):
            for h in self._key_to_hooks[k]:
                h.reset_state()

    def resolve_hooks(self, key: Optional[str] = None) -> None:
        if key is not None:
            self._check_key(key)
        for k in ([key] if key else list(self._key_to_hooks)):
            own = [h for h in self._key_to_hooks[k] if h not in self._shared_hooks]
            self._key_to_hooks[k] = self._topological_sort_hooks(self._shared_hooks + own)
            self._dirty[k] = False

    # -- validation against a model's declared needs (:228-371) --------------------------------
    def validate_requirement(self, module: Any, key: Optional[str] = None) -> None:
        if not (callable(module) and hasattr(module, 'requires')):
            raise BadEncoderProtocolError(
                f'Cannot validate {type(module).__name__}: must implement __call__(self, batch, '
                '*args, **kwargs) and have `requires` attribute')
        if key is not None:
            self._check_key(key)
        for k in ([key] if key is not None else list(self._key_to_hooks)):
            hooks = self._key_to_hooks[k] + self._shared_hooks
            produced = CORE_ATTRIBUTE.union(*(h.produces for h in hooks))
            missing = set(module.requires) - produced
            if missing:
                raise UnresolvableHookDependenciesError(self._suggest(missing, k))

    @staticmethod
    def _suggest(missing: Set[str], key: str) -> str:
        # `list_hooks` is resolved through the module namespace at call time (patchable)
        msg = (f'Cannot resolve the following requirements {missing} from any hook registered '
               f"under key '{key}'.\nSuggestions:")
        for attr in missing:
            hit = False
            for cls in list_hooks():
                produced = getattr(cls, '_cls_produces', set())
                close = difflib.get_close_matches(attr, produced, n=2, cutoff=0.6)
                if attr in produced:
                    msg += (f"\n\t- '{attr}': Found hook that produces '{attr}'. To resolve this, "
                            f"please register '{cls.__name__}' with key '{key}'")
                    hit = True
                elif close:
                    names = ' or '.join(f"'{c}'" for c in close)
                    msg += (f"\n\t- '{attr}': Do you mean {names}?. If so, please update the "
                            f"module requirement with the correct name and register "
                            f"'{cls.__name__}' with key '{key}' to resolve this.")
                    hit = True
                elif attr.lower() in (cls.__doc__ or '').lower():
                    msg += (f"\n\t- '{attr}': Found keyword '{attr}' in '{cls.__name__}' "
                            f"documentation. If this hook produces what you are looking for, "
                            f"update the module requirement with the correct name and register "
                            f"'{cls.__name__}' with key '{key}'.")
                    hit = True
            if not hit:
                msg += (f"\n\t- '{attr}': Can not find any existing hooks that satisfy this "
                        'requirement.')
        return msg

    # -- internals --------------------------------------------------------------------------
    def _check_hook(self, hook: Any) -> None:
        if not isinstance(hook, DGHook):
            raise BadHookProtocolError(
                f'Cannot register hook {type(hook).__name__}: must implement __call__(dg: DGraph, '
                'batch: DGBatch) -> DGBatch, reset_state(), requires and produces properties.')

    def _check_inactive(self) -> None:
        if self._active_key is not None:
            raise RuntimeError('Cannot register hooks while a key is active. Register hooks '
                               'before using `activate`.')

    def _check_key(self, key: str) -> None:
        if key not in self._key_to_hooks:
            raise KeyError(f'{key} was not a declared key in the hook manager')

    @staticmethod
    def _topological_sort_hooks(hooks: List[DGHook]) -> List[DGHook]:
        produced = CORE_ATTRIBUTE.union(*(h.produces for h in hooks))
        missing: Set[str] = set()
        for h in hooks:
            missing |= h.requires - produced
        if missing:
            raise UnresolvableHookDependenciesError(
                'Cannot resolve hook dependencies: required attributes not produced by any '
                f'hook: {missing}')
        n = len(hooks)
        succ: List[List[int]] = [[] for _ in range(n)]
        indeg = [0] * n
        for i, a in enumerate(hooks):
            for j, b in enumerate(hooks):
                if i == j:
                    continue
                if a.produces & b.requires:
                    succ[i].append(j)
                    indeg[j] += 1
                # negatives must exist before neighbours are sampled for them (:420-430); the
                # reference adds this edge on top of a data edge, so it may count twice
                if 'neg' in a.produces and 'nbr_nids' in b.produces:
                    succ[i].append(j)
                    indeg[j] += 1
        ready = deque(i for i in range(n) if indeg[i] == 0)
        order: List[int] = []
        while ready:
            u = ready.popleft()
            order.append(u)
            for v in succ[u]:
                indeg[v] -= 1
                if indeg[v] == 0:
                    ready.append(v)
        if len(order) != n:
            done = set().union(*[hooks[i].produces for i in order]) if order else set()
            msg = 'Cannot resolve hook dependencies:\n'
            for i in range(n):
                if i not in order:
                    msg += (f'\n - {hooks[i]!r} requires {hooks[i].requires - done} but not '
                            'produced (or stuck in cycle)')
            raise UnresolvableHookDependenciesError(msg)
        return [hooks[i] for i in order]
