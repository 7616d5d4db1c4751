"""Hook protocol and base classes.  Same contract as tgm/hooks/base.py:10-103: a hook is any
object with `has_state`, `requires`, `produces`, `__call__(dg, batch) -> batch` and
`reset_state()`; class-level `_cls_requires/_cls_produces` are read from the LEAF class
(:41-43), `id` suffixes every produced attribute (:46-51,:69-76), `seed_keys` join `requires`
(:101-103)."""
from __future__ import annotations

from abc import ABC, abstractmethod
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


class BaseDGHook(ABC):
    """Bookkeeping shared by every hook, with the construction protocol of the reference's
    dataclass bases (tgm/hooks/base.py:27-76) so that user hooks written against them run
    unchanged: `super().__init__(_requires=..., _produces=..., _id=..., has_state=...)` and an
    idempotent `__post_init__()` that folds the LEAF class's `_cls_requires/_cls_produces` into the
    instance sets (:41-43).  As upstream, `__init__` stores `has_state` on the instance (default
    False, whatever the class attribute says), so a stateful user hook sets it itself.
    The hooks of this package call `_init_hook` instead and keep the class attribute."""
    has_state: bool = False
    _cls_requires: Set[str] = set()
    _cls_produces: Set[str] = set()

    def __init__(self, _requires: Optional[Set[str]] = None, _produces: Optional[Set[str]] = None,
                 _id: Optional[str] = None, has_state: bool = False) -> None:
        self._requires: Set[str] = set() if _requires is None else _requires
        self._produces: Set[str] = set() if _produces is None else _produces
        self._id = _id
        self.has_state = has_state
        self.__post_init__()

    def __post_init__(self) -> None:
        leaf = type(self).__dict__
        self._requires.update(leaf.get('_cls_requires', ()))
        self._produces.update(leaf.get('_cls_produces', ()))

    def _init_hook(self, id: Optional[str] = None, seed_keys: Optional[List[str]] = None) -> None:
        self._requires, self._produces, self._id = set(), set(), id
        BaseDGHook.__post_init__(self)
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

    @abstractmethod
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
    """Hooks that take extra seed attribute names: `seed_keys` join `requires` (:92-103)."""

    def __init__(self, _requires: Optional[Set[str]] = None, _produces: Optional[Set[str]] = None,
                 _id: Optional[str] = None, has_state: bool = False,
                 seed_keys: Optional[List[str]] = None) -> None:
        self.seed_keys = [] if seed_keys is None else seed_keys
        super().__init__(_requires, _produces, _id, has_state)

    def __post_init__(self) -> None:
        super().__post_init__()
        self._requires.update(getattr(self, 'seed_keys', None) or [])
