"""Import paths of the reference's encoder package (tgm/nn/encoder/__init__.py) for the encoders on
the hot path; the implementations live in tgm_b200/nn/{tgat,dygformer,tgn}.py."""
from tgm_b200.nn.dygformer import DyGFormer
from tgm_b200.nn.tgat import TGAT
from tgm_b200.nn.tgn import (GraphAttentionEmbedding, IdentityMessage, LastAggregator,
                             MeanAggregator, TGNMemory)

__all__ = ['DyGFormer', 'TGAT', 'GraphAttentionEmbedding', 'IdentityMessage', 'LastAggregator',
           'MeanAggregator', 'TGNMemory']
