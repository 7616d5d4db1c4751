"""Aggregation modules of the hot path (tgm/nn of the reference, the parts SURVEY.md section 8
puts on the path): `torch.nn.Module`s with the reference's parameter names and shapes, so
`state_dict`s interchange, executing on the CUDA library (include/tgm_b200.h); differentiable with
dropout 0 (`tgm_attn_backward`, `tgm_dyg_backward`, `tgm_tgn_backward`, `tgm_gae_backward`)."""
from .attention import MergeLayer, TemporalAttention, Time2Vec, masked_mean
from .dygformer import DyGFormer
from .tgat import TGAT
from .tgn import (GraphAttentionEmbedding, IdentityMessage, LastAggregator, MeanAggregator,
                  TGNMemory, TransformerConv)

__all__ = ['TemporalAttention', 'Time2Vec', 'MergeLayer', 'TGAT', 'DyGFormer', 'masked_mean',
           'TGNMemory', 'IdentityMessage', 'LastAggregator', 'MeanAggregator', 'GraphAttentionEmbedding',
           'TransformerConv']
