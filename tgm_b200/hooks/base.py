"""Hook protocol and base classes.  Same contract as tgm/hooks/base.py:10-103: a hook is any
object with `has_state`, `requires`, `produces`, `__call__(dg, batch) -> batch` and
`reset_state()`; class-level `_cls_requires/_cls_produces` are read from the LEAF class
(:41-43), `id` suffixes every produced attribute (:46-51,:69-76), `seed_keys` join `requires`
(:101-103)."""
from __future__ import annotations

from typing import Any, List, Optional, Protocol, Set, runtime_checkable


@runtime_checkable
class DGHook(Protocol):
    has_state: bool

    @property
    def requires(self) -> Set[str]: ...

    @property
    def produces(self) -> Set[str]: ...

    def __call__(self, dg, batch): ...

    def reset_state(self) -> None: ...


class BaseDGHook:
    """Common bookkeeping; subclasses call `_init_hook` at the end of their __init__."""
    has_state: bool = False
    _cls_requires: Set[str] = set()
    _cls_produces: Set[str] = set()

    def _init_hook(self, id: Optional[str] = None, seed_keys: Optional[List[str]] = None) -> None:
        leaf = type(self).__dict__
        self._requires: Set[str] = set(leaf.get('_cls_requires', set()))
        self._produces: Set[str] = set(leaf.get('_cls_produces', set()))
        self._id = id
        self.seed_keys = seed_keys
        if seed_keys:
            self._requires.update(seed_keys)

    @property
    def requires(self) -> Set[str]:
        return self._requires

    @property
    def produces(self) -> Set[str]:
        if self._id is None:
            return self._produces
        return {f'{name}_{self._id}' for name in self._produces}

    def __repr__(self) -> str:
        name = type(self).__name__
        return f'{name}_{self._id}' if self._id else name

    def __call__(self, dg, batch):
        raise NotImplementedError

    def reset_state(self) -> None:
        pass

    def add_batch_attribute(self, batch, name: str, value: Any) -> None:
        setattr(batch, f'{name}_{self._id}' if self._id else name, value)


class StatelessHook(BaseDGHook):
    has_state = False


class StatefulHook(BaseDGHook):
    has_state = True


class SeedableHook(BaseDGHook):
    """Marker for hooks that take extra seed attribute names (`seed_keys`)."""
