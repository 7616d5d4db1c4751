"""tgm_b200: B200-native temporal neighbor sampling + aggregation behind the TGM API.

Public surface mirrors tgm-team/tgm for the hot path only: DGData ->
DGraph (device-resident store) -> DGDataLoader -> HookManager -> RecencyNeighborHook, plus the
stateless windowed sampler `RecencyCSR` and the aggregation ops in `tgm_b200.nn`.
The compute path is the CUDA library `tgm_b200/csrc/libtgm_b200.so` (C ABI: include/tgm_b200.h);
there is no CPU fallback.
"""
from . import _cabi  # noqa: F401  (fails loudly when the CUDA library is missing)
from .constants import PADDED_NODE_ID
from .core import DGBatch, DGraph, TimeDeltaDG
from .data import DGData, DGDataLoader
from .hooks import (DeduplicationHook, HookManager, NeighborSamplerHook,
                    RandomNegativeEdgeSamplerHook, RecencyNeighborHook, TGBNegativeEdgeSamplerHook,
                    TGBTHGNegativeEdgeSamplerHook, TGBTKGNegativeEdgeSamplerHook)
from .sampler import RecencyCSR

__version__ = '0.1.0'
__all__ = ['DGraph', 'DGBatch', 'DGData', 'DGDataLoader', 'TimeDeltaDG', 'HookManager',
           'RecencyNeighborHook', 'NeighborSamplerHook', 'RandomNegativeEdgeSamplerHook', 'DeduplicationHook',
           'TGBNegativeEdgeSamplerHook', 'TGBTHGNegativeEdgeSamplerHook', 'TGBTKGNegativeEdgeSamplerHook',
           'RecencyCSR', 'PADDED_NODE_ID']
