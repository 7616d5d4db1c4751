"""TGAT encoder on the B200 library: same constructor, parameter names and forward signature
as tgm/nn/encoder/tgat.py:41-149 (state_dicts interchange with the reference)."""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn
from torch import Tensor

from tgm_b200.nn.attention import MergeLayer, TemporalAttention, Time2Vec, gather_rows


class TGAT(nn.Module):
    """Temporal Graph Attention Network.  Differentiable end to end (attention layers through
    tgm_attn_backward, merge layers through torch's Linear) with dropout 0."""

    def __init__(self, node_dim: int, edge_dim: int, time_dim: int, embed_dim: int,
                 num_layers: int, n_heads: int = 2, dropout: float = 0.1) -> None:
        super().__init__()
        self.num_layers, self.embed_dim = num_layers, embed_dim
        self.time_encoder = Time2Vec(time_dim=time_dim)
        self.attn, self.merge_layers = nn.ModuleList(), nn.ModuleList()
        for i in range(num_layers):
            self.attn.append(TemporalAttention(
                n_heads=n_heads, node_dim=node_dim if i == 0 else embed_dim, edge_dim=edge_dim,
                time_dim=time_dim, dropout=dropout))
            self.merge_layers.append(MergeLayer(
                in_dim1=self.attn[-1].out_dim, in_dim2=node_dim, hidden_dim=embed_dim,
                output_dim=embed_dim))

    def forward(self, node_x: Tensor, seed_nids: List[Tensor], seed_times: List[Tensor],
                nbr_nids: List[Tensor], nbr_edge_x: List[Tensor],
                nbr_edge_time: List[Tensor]) -> Tensor:
        """Hop recursion of tgat.py:122-149; z[j][i] = embedding of hop-i nodes after j layers."""
        L = self.num_layers
        z: Dict[int, Dict[int, Tensor]] = {j: {} for j in range(L + 1)}
        z[0][0] = gather_rows(node_x, seed_nids[0])
        for i in range(1, L + 1):
            z[0][i] = gather_rows(node_x, nbr_nids[i - 1].flatten())
        for j in range(1, L + 1):
            for i in range(L - j + 1):
                n = z[j - 1][i].size(0)
                k = nbr_nids[j - 1].shape[-1]  # tgat.py:139 (all hops share k)
                out = self.attn[j - 1].forward_fused(
                    self.time_encoder, node_x=z[j - 1][i],
                    nbr_node_feat=z[j - 1][i + 1].reshape(n, k, -1), edge_feat=nbr_edge_x[i],
                    seed_times=seed_times[i], nbr_times=nbr_edge_time[i], nbr_nids=nbr_nids[i])
                z[j][i] = self.merge_layers[j - 1](out, z[0][i])
        return z[L][0]
