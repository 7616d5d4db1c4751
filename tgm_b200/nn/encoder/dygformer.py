"""`tgm.nn.encoder.dygformer` import path: see tgm_b200/nn/dygformer.py."""
from tgm_b200.nn.dygformer import *  # noqa: F401,F403
from tgm_b200.nn import dygformer as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
