"""The storage plug-in for an UNMODIFIED tgm-team/tgm installation (INTEGRATION.md section 2 as
code): `install()` registers the B200 event store as a backend of the reference's own registry,
so the reference's `DGraph`, loader and hooks run on it.

Reference interfaces this binds to (tgm-team/tgm @ 5183dc9):
  tgm/core/_storage/base.py:20-118          DGStorageBase (the ABC every backend implements)
  tgm/core/_storage/backends/__init__.py:3-7 DGStorageBackends (name -> class), DGStorage
  tgm/core/_storage/__init__.py:18-28       set_dg_storage_backend
  tgm/core/graph.py:13                       `from tgm.core._storage import DGStorage` (a by-value
                                             import: the name has to be rebound there too)

Nothing here computes: the class is `DeviceCOOStorage` (tgm_b200/core/storage.py) under the
reference's ABC.  The reference constructs a backend as `DGStorage(data)`; the device it should live
on is fixed at install time (default: the current CUDA device; without one the store is
metadata-only -- slice bounds, counts, times -- and every edge getter raises the no-CPU-fallback
error).
"""
from __future__ import annotations

from typing import Optional

import torch

from tgm_b200.core.storage import DeviceCOOStorage

BACKEND_NAME = 'B200Storage'


def make_backend(device: 'torch.device | str | None' = 'auto'):
    """The backend class for the installed `tgm` package: DeviceCOOStorage's getters under
    `tgm.core._storage.base.DGStorageBase`, constructed as `cls(data)`."""
    from tgm.core._storage.base import DGStorageBase  # the reference's ABC

    if device == 'auto':
        device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else None
    fixed: Optional[torch.device] = None if device is None else torch.device(device)

    class B200Storage(DeviceCOOStorage, DGStorageBase):  # type: ignore[misc]
        __doc__ = DeviceCOOStorage.__doc__

        def __init__(self, data, device: 'torch.device | str | None' = None) -> None:
            DeviceCOOStorage.__init__(self, data, device=fixed if device is None else device)

    return B200Storage


def install(device: 'torch.device | str | None' = 'auto', make_default: bool = True):
    """Register the backend with the installed `tgm`; with `make_default` it also becomes what
    `DGraph(data)` constructs.  Returns the class.  `uninstall()` restores the array backend."""
    import tgm.core._storage as storage_pkg
    import tgm.core._storage.backends as backends
    import tgm.core.graph as graph

    cls = make_backend(device)
    backends.DGStorageBackends[BACKEND_NAME] = cls
    if make_default:
        storage_pkg.set_dg_storage_backend(cls)
        backends.DGStorage = cls
        graph.DGStorage = cls  # graph.py:13 imported the class by value
    return cls


def uninstall() -> None:
    import tgm.core._storage as storage_pkg
    import tgm.core._storage.backends as backends
    import tgm.core.graph as graph

    backends.DGStorageBackends.pop(BACKEND_NAME, None)
    default = backends.DGStorageBackends['ArrayBackend']
    storage_pkg.set_dg_storage_backend(default)
    backends.DGStorage = default
    graph.DGStorage = default
