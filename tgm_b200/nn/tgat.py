"""TGAT encoder on the B200 library: same constructor, parameter names and forward signature
as tgm/nn/encoder/tgat.py:41-149 (state_dicts interchange with the reference)."""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn
from torch import Tensor

import ctypes

from tgm_b200 import _cabi
from tgm_b200.nn.attention import (MergeLayer, TemporalAttention, Time2Vec, _NativeHandle, _f32,
                                   gather_rows)


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
        self._native = _NativeHandle(_cabi.lib.tgm_tgat_destroy)

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
        if self.num_layers > 4 or not node_x.is_cuda:
            return False
        if torch.is_grad_enabled() and (node_x.requires_grad or
                                        any(p.requires_grad for p in self.parameters())):
            return False
        k = nbr_nids[0].shape[-1]
        if any(n.shape[-1] != k for n in nbr_nids[:self.num_layers]):
            return False
        rows = nbr_nids[0].numel() // k  # the hop recursion: every slot of hop i seeds hop i+1
        for n in nbr_nids[:self.num_layers]:
            if n.numel() != rows * k:
                return False
            rows *= k
        return all(a.covers_hops(self.time_encoder, k, node_x.device) for a in self.attn)

    def _forward_hops_batched(self, node_x, seed_nids, seed_times, nbr_nids, nbr_edge_x,
                              nbr_edge_time) -> Tensor:
        """The whole recursion as ONE native call (tgm_tgat_forward, csrc/tgat.cu).  Hop i+1's
        nodes are hop i's neighbour slots, so with the hops stored back to back the neighbour
        features of hops 0..m are simply a later row range of the same buffer: layer j is one
        folded attention chain + one merge layer over the rows of hops 0..L-j, and the per-hop id /
        time / edge-feature arrays are read where the sampler wrote them (no concatenation)."""
        from tgm_b200.sampler import LazyEdgeRows
        L = self.num_layers
        k = nbr_nids[0].shape[-1]
        dev = node_x.device
        attn_h = [a._handle(self.time_encoder, dev) for a in self.attn]
        merge_h = [m._handle(dev) for m in self.merge_layers]
        ver = tuple(h.value for h in attn_h + merge_h)
        if self._native.version != ver:
            self._native.free()
            _cabi.check(_cabi.lib.tgm_tgat_create(
                ctypes.byref(self._native.h), L, (ctypes.c_void_p * L)(*attn_h),
                (ctypes.c_void_p * L)(*merge_h), dev.index))
            self._native.version = ver
        i32 = lambda t: t.to(device=dev, dtype=torch.int32).contiguous()  # noqa: E731
        i64 = lambda t: t.to(device=dev, dtype=torch.int64).contiguous()  # noqa: E731
        seeds = i32(seed_nids[0].reshape(-1))
        nids = [i32(nbr_nids[i].reshape(-1, k)) for i in range(L)]
        st = [i64(seed_times[i].reshape(-1)) for i in range(L)]
        nt = [i64(nbr_edge_time[i].reshape(-1, k)) for i in range(L)]
        rows = seeds.numel()
        for i in range(L):
            if nids[i].shape[0] != rows or st[i].numel() != rows or nt[i].shape[0] != rows:
                raise ValueError(f'TGAT: hop {i} arrays do not hold {rows} seeds')
            rows *= k
        ptrs = lambda ts: (ctypes.c_void_p * L)(*[t.data_ptr() for t in ts])  # noqa: E731
        x = _f32(node_x)
        keep = [x, seeds, nids, st, nt]
        lazy = all(isinstance(e, LazyEdgeRows) for e in nbr_edge_x[:L])
        t0 = nbr_edge_x[0].table if lazy else None
        if lazy and all(e.table.data_ptr() == t0.data_ptr() and e.table.shape == t0.shape and
                        e.table.dtype == t0.dtype for e in nbr_edge_x[:L]):
            table = _f32(nbr_edge_x[0].table)
            er = [i32(e.rows.reshape(-1, k)) for e in nbr_edge_x[:L]]
            keep += [table, er]
            edge_args = (None, table.data_ptr(), ptrs(er))
        else:
            ef = [_f32(e.materialize() if isinstance(e, LazyEdgeRows) else e)
                  for e in nbr_edge_x[:L]]
            keep.append(ef)
            edge_args = (ptrs(ef), None, None)
        out = torch.empty((seeds.numel(), self.embed_dim), dtype=torch.float32, device=dev)
        _cabi.check(_cabi.lib.tgm_tgat_forward(
            self._native.h, x.data_ptr(), x.shape[0], seeds.data_ptr(), seeds.numel(), ptrs(nids),
            ptrs(st), ptrs(nt), *edge_args, k, out.data_ptr(), _cabi.current_stream(dev)))
        del keep
        return out
