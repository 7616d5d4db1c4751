from .timedelta import TimeDeltaDG
from .batch import DGBatch
from .storage import (DGSliceTracker, DGStorageBase, DeviceCOOStorage, DGStorageBackends,
                      get_dg_storage_backend, set_dg_storage_backend)
from .graph import DGraph
