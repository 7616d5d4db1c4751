"""Import paths of the reference's module package (tgm/nn/modules/__init__.py) for the modules on
the hot path; the implementations live in tgm_b200/nn/attention.py."""
from tgm_b200.nn.attention import TemporalAttention, Time2Vec

__all__ = ['TemporalAttention', 'Time2Vec']
