from .base import BaseDGHook, DGHook, SeedableHook, StatefulHook, StatelessHook
from .hook_manager import HookManager
from .recency import RecencyNeighborHook
from .negatives import RandomNegativeEdgeSamplerHook
from .tgb_negatives import (TGBNegativeEdgeSamplerBase, TGBNegativeEdgeSamplerHook,
                            TGBTHGNegativeEdgeSamplerHook, TGBTKGNegativeEdgeSamplerHook)
from .dedup import DeduplicationHook
from .uniform import NeighborSamplerHook
from .registry import hook, list_hooks
