"""`tgm.nn.encoder.tgat` import path: see tgm_b200/nn/tgat.py."""
from tgm_b200.nn.tgat import *  # noqa: F401,F403
from tgm_b200.nn import tgat as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
