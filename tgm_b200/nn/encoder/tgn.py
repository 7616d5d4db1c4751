"""`tgm.nn.encoder.tgn` import path: see tgm_b200/nn/tgn.py."""
from tgm_b200.nn.tgn import *  # noqa: F401,F403
from tgm_b200.nn import tgn as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
