"""DyGFormer on the B200 library: same constructor, parameter names and forward signature as
tgm/nn/encoder/dygformer.py:146-444 (state_dicts interchange with the reference); the forward
pass runs in `tgm_dyg_forward` (include/tgm_b200.h).  Differentiable with respect to every
parameter when autograd is recording (`tgm_dyg_backward` behind a torch.autograd.Function; dropout
must be 0 in training mode -- a dropout mask cannot follow the reference's RNG stream)."""
from __future__ import annotations

import ctypes
from typing import Callable, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from tgm_b200 import _cabi
from tgm_b200.nn.attention import Time2Vec, _NativeHandle, _f32, _need_cuda, _version


class _DyGFunction(torch.autograd.Function):
    """tgm_dyg_forward / tgm_dyg_backward as one differentiable op.  The parameters are passed
    explicitly so autograd routes their gradients; backward recomputes the forward activations
    inside the library, so only the inputs are kept."""

    @staticmethod
    def forward(ctx, module, table, src, dst, t, nb, nt, nx, *params):
        zs, zd = module._run_forward(table, src, dst, t, nb, nt, nx)
        ctx.module, ctx.inputs = module, (table, src, dst, t, nb, nt, nx)
        return zs, zd

    @staticmethod
    def backward(ctx, d_zs, d_zd):
        module = ctx.module
        table, src, dst, t, nb, nt, nx = ctx.inputs
        dev, B = table.device, src.numel()
        z = lambda d: (torch.zeros((B, module.output_dim), dtype=torch.float32, device=dev)
                       if d is None else _f32(d))
        d_zs, d_zd = z(d_zs), z(d_zd)
        params = list(module.parameters())
        flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
        grads = [v.view_as(p) for v, p in zip(flat.split([p.numel() for p in params]), params)]
        by_param = {id(p): g for p, g in zip(params, grads)}
        G = lambda p: by_param[id(p)].data_ptr()
        layers = (_cabi.DygLayer * max(1, module.num_layers))()
        gr = _cabi.DygGrads()
        module._fill_tables(gr, layers, G)
        _cabi.check(_cabi.lib.tgm_dyg_backward(
            module._handle(dev), table.data_ptr(), table.shape[0], src.data_ptr(), dst.data_ptr(),
            t.data_ptr(), nb.data_ptr(), nt.data_ptr(), nx.data_ptr(), B, d_zs.data_ptr(),
            d_zd.data_ptr(), ctypes.byref(gr), _cabi.current_stream(dev)))
        return (None,) * 8 + tuple(grads)


class NeighborCooccurrenceEncoder(nn.Module):
    def __init__(self, feat_dim: int, device: str = 'cpu') -> None:
        super().__init__()
        self.feat_dim = feat_dim
        self.neighbor_co_occurrence_encoder = nn.Sequential(
            nn.Linear(1, feat_dim), nn.ReLU(), nn.Linear(feat_dim, feat_dim))


class TransformerEncoder(nn.Module):
    def __init__(self, attention_dim: int, num_heads: int, dropout: float = 0.1) -> None:
        super().__init__()
        self.attention_dim, self.num_heads = attention_dim, num_heads
        self.multi_head_attention = nn.MultiheadAttention(attention_dim, num_heads, dropout=dropout)
        self.dropout = nn.Dropout(dropout)
        self.linear_layers = nn.ModuleList([nn.Linear(attention_dim, 4 * attention_dim),
                                            nn.Linear(4 * attention_dim, attention_dim)])
        self.norm_layers = nn.ModuleList([nn.LayerNorm(attention_dim), nn.LayerNorm(attention_dim)])


class DyGFormer(nn.Module):
    def __init__(self, node_feat_dim: int, edge_x_dim: int, time_feat_dim: int,
                 channel_embedding_dim: int, output_dim: int = 172, patch_size: int = 1,
                 num_layers: int = 2, num_heads: int = 2, dropout: float = 0.1,
                 max_input_sequence_length: int = 512, num_channels: int = 4,
                 time_encoder: Callable[..., nn.Module] = Time2Vec, device: str = 'cpu') -> None:
        super().__init__()
        if max_input_sequence_length % patch_size != 0:
            raise ValueError('Max sequence length must be a multiple of path size')
        if num_channels != 4:
            raise ValueError('DyGFormer stacks exactly 4 channels (dygformer.py:401-409)')
        self.node_feat_dim, self.edge_x_dim, self.time_feat_dim = node_feat_dim, edge_x_dim, time_feat_dim
        self.channel_embedding_dim, self.patch_size = channel_embedding_dim, patch_size
        self.max_input_sequence_length = max_input_sequence_length
        self.num_channels, self.output_dim = num_channels, output_dim
        self.num_patches = max_input_sequence_length // patch_size
        self.num_heads, self.num_layers = num_heads, num_layers
        C = channel_embedding_dim
        self.time_encoder = time_encoder(time_feat_dim)
        self.co_occurrence_encoder = NeighborCooccurrenceEncoder(C)
        self.projection_layer = nn.ModuleDict({
            'node': nn.Linear(patch_size * node_feat_dim, C),
            'edge': nn.Linear(patch_size * edge_x_dim, C),
            'time': nn.Linear(patch_size * time_feat_dim, C),
            'neighbor_co_occurrence': nn.Linear(patch_size * C, C)})
        self.transformers = nn.ModuleList(
            [TransformerEncoder(num_channels * C, num_heads, dropout) for _ in range(num_layers)])
        self.output_layer = nn.Linear(num_channels * C, output_dim)
        self._native = _NativeHandle(_cabi.lib.tgm_dyg_destroy)
        self.to(device)

    def _fill_tables(self, table, layers, P) -> None:
        """Pointer tables of include/tgm_b200.h (tgm_dyg_params / tgm_dyg_grads share the field
        order); `P(tensor)` yields the device pointer to put in the slot of that parameter."""
        for i, tr in enumerate(self.transformers):
            mha, ly = tr.multi_head_attention, layers[i]
            ly.in_proj_w, ly.in_proj_b = P(mha.in_proj_weight), P(mha.in_proj_bias)
            ly.out_proj_w, ly.out_proj_b = P(mha.out_proj.weight), P(mha.out_proj.bias)
            ly.ffn1_w, ly.ffn1_b = P(tr.linear_layers[0].weight), P(tr.linear_layers[0].bias)
            ly.ffn2_w, ly.ffn2_b = P(tr.linear_layers[1].weight), P(tr.linear_layers[1].bias)
            ly.ln0_w, ly.ln0_b = P(tr.norm_layers[0].weight), P(tr.norm_layers[0].bias)
            ly.ln1_w, ly.ln1_b = P(tr.norm_layers[1].weight), P(tr.norm_layers[1].bias)
        table.t2v_w, table.t2v_b = P(self.time_encoder.w.weight), P(self.time_encoder.w.bias)
        mlp = self.co_occurrence_encoder.neighbor_co_occurrence_encoder
        table.cooc_w1, table.cooc_b1 = P(mlp[0].weight), P(mlp[0].bias)
        table.cooc_w2, table.cooc_b2 = P(mlp[2].weight), P(mlp[2].bias)
        for c, name in enumerate(('node', 'edge', 'time', 'neighbor_co_occurrence')):
            table.proj_w[c] = P(self.projection_layer[name].weight)
            table.proj_b[c] = P(self.projection_layer[name].bias)
        table.layers = ctypes.cast(layers, ctypes.POINTER(_cabi.DygLayer))
        table.out_w, table.out_b = P(self.output_layer.weight), P(self.output_layer.bias)

    def _handle(self, dev: torch.device) -> ctypes.c_void_p:
        params = list(self.parameters())
        ver = _version(params)
        if self._native.version == ver:
            return self._native.h
        for p in params:
            _need_cuda(p, 'DyGFormer parameters')
        keep = []  # contiguous fp32 copies must outlive the call

        def P(t: Tensor) -> int:
            c = _f32(t)
            keep.append(c)
            return c.data_ptr()

        layers = (_cabi.DygLayer * max(1, self.num_layers))()
        pr = _cabi.DygParams()
        pr.node_dim, pr.edge_dim, pr.time_dim = self.node_feat_dim, self.edge_x_dim, self.time_feat_dim
        pr.channel_dim, pr.out_dim, pr.patch_size = self.channel_embedding_dim, self.output_dim, self.patch_size
        pr.num_layers, pr.num_heads = self.num_layers, self.num_heads
        pr.seq_len = self.max_input_sequence_length
        pr.ln_eps = float(self.transformers[0].norm_layers[0].eps) if self.num_layers else 1e-5
        self._fill_tables(pr, layers, P)
        if self._native.h.value and getattr(self, '_native_dev', None) == dev:
            # same shapes, new values (optimizer step): refresh the copies in place
            _cabi.check(_cabi.lib.tgm_dyg_set_params(self._native.h, ctypes.byref(pr),
                                                     _cabi.current_stream(dev)))
        else:
            self._native.free()
            _cabi.check(_cabi.lib.tgm_dyg_create(ctypes.byref(self._native.h), ctypes.byref(pr),
                                                 dev.index))
            self._native_dev = dev
        self._native.version = ver
        return self._native.h

    def _run_forward(self, table, src, dst, t, nb, nt, nx) -> Tuple[Tensor, Tensor]:
        dev, B = table.device, src.numel()
        zs = torch.empty((B, self.output_dim), dtype=torch.float32, device=dev)
        zd = torch.empty((B, self.output_dim), dtype=torch.float32, device=dev)
        _cabi.check(_cabi.lib.tgm_dyg_forward(
            self._handle(dev), table.data_ptr(), table.shape[0], src.data_ptr(), dst.data_ptr(),
            t.data_ptr(), nb.data_ptr(), nt.data_ptr(), nx.data_ptr(), B, zs.data_ptr(),
            zd.data_ptr(), _cabi.current_stream(dev)))
        return zs, zd

    def forward(self, node_x: Tensor, edge_index: Tensor, edge_time: Tensor, neighbours: Tensor,
                neighbours_time: Tensor, neighbours_edge_feat: Tensor) -> Tuple[Tensor, Tensor]:
        """Rows [0, E_b) of `neighbours*` belong to the sources, [E_b, 2 E_b) to the destinations
        (dygformer.py:262-270); extra rows (e.g. negatives) are ignored, as in the reference."""
        if self.training:
            ps = [m.p for m in self.modules() if isinstance(m, nn.Dropout)] + \
                 [tr.multi_head_attention.dropout for tr in self.transformers]
            if any(p > 0 for p in ps):
                from tgm_b200.nn.attention import warn_dropout_disabled
                warn_dropout_disabled('DyGFormer', max(ps))
        dev = _need_cuda(node_x, 'DyGFormer')
        B = edge_index.shape[1]
        k = self.max_input_sequence_length - 1
        if neighbours.shape[1] != k:
            raise ValueError(f'expected {k} sampled neighbours per node (max_input_sequence_length'
                             f' - 1), got {neighbours.shape[1]}')
        src = edge_index[0].to(device=dev, dtype=torch.int32).contiguous()
        dst = edge_index[1].to(device=dev, dtype=torch.int32).contiguous()
        t = edge_time.to(device=dev, dtype=torch.int64).contiguous()
        nb = neighbours[:2 * B].to(device=dev, dtype=torch.int32).contiguous()
        nt = neighbours_time[:2 * B].to(device=dev, dtype=torch.int64).contiguous()
        nx = _f32(neighbours_edge_feat[:2 * B].to(dev))
        table = _f32(node_x)
        params = list(self.parameters())
        if B > 0 and torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _DyGFunction.apply(self, table, src, dst, t, nb, nt, nx, *params)
        with torch.no_grad():
            return self._run_forward(table, src, dst, t, nb, nt, nx)
