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
        if self._can_batch_hops(node_x, nbr_nids):
            with torch.no_grad():
                return self._forward_hops_batched(node_x, seed_nids, seed_times, nbr_nids,
                                                  nbr_edge_x, nbr_edge_time)
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

    def _can_batch_hops(self, node_x: Tensor, nbr_nids: List[Tensor]) -> bool:
        """Inference on the device with one k for all hops (tgat.py:139) and shapes the folded
        attention chain covers: every layer runs once over the rows of all its hops."""
        if self.num_layers < 2 or self.num_layers > 4 or not node_x.is_cuda:
            return False
        if torch.is_grad_enabled() and (node_x.requires_grad or
                                        any(p.requires_grad for p in self.parameters())):
            return False
        k = nbr_nids[0].shape[-1]
        if any(n.shape[-1] != k for n in nbr_nids[:self.num_layers]):
            return False
        return all(a.covers_hops(self.time_encoder, k, node_x.device) for a in self.attn)

    def _forward_hops_batched(self, node_x, seed_nids, seed_times, nbr_nids, nbr_edge_x,
                              nbr_edge_time) -> Tensor:
        """The same recursion with each layer's hops as one row range.  Hop i+1's nodes are hop i's
        neighbour slots, so with the hops stored back to back -- rows [off[i], off[i+1]) -- the
        neighbour features of hops 0..m are simply rows [off[1], off[m+2]): no copies, and layer j
        is one attention call + one merge call over off[L-j+1] rows instead of L-j+1 of each."""
        L = self.num_layers
        k = nbr_nids[0].shape[-1]
        dev = node_x.device
        ids = [seed_nids[0].reshape(-1)] + [nbr_nids[i].reshape(-1) for i in range(L)]
        off = [0]
        for t in ids:
            off.append(off[-1] + t.numel())
        z0 = gather_rows(node_x, torch.cat([t.to(device=dev, dtype=torch.int32) for t in ids]))
        st = torch.cat([seed_times[i].reshape(-1).to(torch.int64) for i in range(L)])
        nt = torch.cat([nbr_edge_time[i].reshape(-1, k).to(torch.int64) for i in range(L)])
        ni = torch.cat([nbr_nids[i].reshape(-1, k).to(torch.int32) for i in range(L)])
        prev = z0
        for j in range(1, L + 1):
            rows = off[L - j + 1]  # seeds of hops 0 .. L-j
            out = self.attn[j - 1].forward_hops(
                self.time_encoder, prev[:rows], prev[off[1]:off[L - j + 2]].reshape(rows, k, -1),
                [nbr_edge_x[i] for i in range(L - j + 1)], st[:rows], nt[:rows], ni[:rows])
            prev = self.merge_layers[j - 1](out, z0[:rows])
        return prev[:off[1]]
