"""Time2Vec, TemporalAttention and MergeLayer on the B200 library.

Same constructors, parameter names/shapes and forward signatures as the reference
(tgm/nn/modules/time_encoding.py:6-24, tgm/nn/modules/attention.py:5-128,
tgm/nn/encoder/tgat.py:11-38).  The computation runs in `tgm_attn_forward` / `tgm_mlp2_forward` /
`tgm_time2vec`; `TemporalAttention.forward_fused` is differentiable (`tgm_attn_backward` behind a
torch.autograd.Function) for every parameter, the seed features and the neighbour features, with
dropout 0 (a dropout mask cannot match the reference's RNG stream anyway).  There is no CPU path:
parameters must live on a CUDA device when forward is called.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from tgm_b200 import _cabi


def _need_cuda(t: Tensor, what: str) -> torch.device:
    if not t.is_cuda:
        raise _cabi.TGMNativeError(-3, f'{what} needs CUDA tensors (tgm_b200 has no CPU fallback)')
    return t.device


def _f32(t: Tensor) -> Tensor:
    return t.detach().to(torch.float32).contiguous()


class _NativeHandle:
    """Owns a C handle created from the module's current parameters; rebuilt when they change."""

    def __init__(self, destroy) -> None:
        self.h = ctypes.c_void_p()
        self._destroy = destroy
        self.version = None

    def free(self) -> None:
        if self.h.value:
            self._destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self) -> None:
        try:
            self.free()
        except Exception:  # noqa: BLE001  interpreter shutdown
            pass


def _version(params) -> tuple:
    return tuple((p.data_ptr(), p._version, str(p.device)) for p in params)


class _FusedAttention(torch.autograd.Function):
    """tgm_attn_forward / tgm_attn_backward as one differentiable op.  The parameters are passed
    explicitly so autograd routes their gradients; the C handle recomputes the forward
    intermediates in backward, so nothing but the inputs is saved."""

    @staticmethod
    def forward(ctx, module, te, seed_times, nbr_times, nbr_nids, node_x, nbr_node_feat, edge_feat,
                *params):
        dev = node_x.device
        S, k = nbr_nids.shape
        args = [_f32(node_x), _f32(nbr_node_feat), _f32(edge_feat),
                seed_times.to(torch.int64).contiguous(), nbr_times.to(torch.int64).contiguous(),
                nbr_nids.to(torch.int32).contiguous()]
        out = torch.empty((S, module.out_dim), dtype=torch.float32, device=dev)
        _cabi.check(_cabi.lib.tgm_attn_forward(
            module._handle(te, dev), *[a.data_ptr() for a in args], S, k, out.data_ptr(),
            _cabi.current_stream(dev)))
        ctx.module, ctx.te, ctx.args = module, te, args
        ctx.need = (node_x.requires_grad, nbr_node_feat.requires_grad, edge_feat.requires_grad)
        return out

    @staticmethod
    def backward(ctx, d_out):
        module, te, args = ctx.module, ctx.te, ctx.args
        dev = d_out.device
        S, k = args[5].shape
        d_out = _f32(d_out)
        od, key, td = module.out_dim, module.node_dim + module.edge_dim + module.time_dim, module.time_dim
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        d_x = z(S, module.node_dim)
        d_nbr = z(S, k, module.node_dim) if ctx.need[1] else None
        d_edge = z(S, k, module.edge_dim) if ctx.need[2] else None
        gp = [z(od, od), z(2 * od, key), z(od, od), z(od), z(od), z(od), z(td), z(td)]
        _cabi.check(_cabi.lib.tgm_attn_backward(
            module._handle(te, dev), *[a.data_ptr() for a in args], S, k, d_out.data_ptr(),
            d_x.data_ptr(), _cabi.ptr(d_nbr), _cabi.ptr(d_edge), *[g.data_ptr() for g in gp],
            _cabi.current_stream(dev)))
        gp[6] = gp[6].reshape(td, 1)  # Time2Vec weight is Linear(1, d).weight
        return (None, None, None, None, None, d_x if ctx.need[0] else None, d_nbr, d_edge, *gp)


class Time2Vec(nn.Module):
    """cos(Linear(1, time_dim)(t)), w initialised to 1/10^linspace(0,9,d), b = 0."""

    def __init__(self, time_dim: int) -> None:
        super().__init__()
        self.time_dim = time_dim
        self.w = nn.Linear(1, time_dim)
        w = (1 / 10 ** np.linspace(0, 9, time_dim)).reshape(time_dim, 1)
        self.w.weight = nn.Parameter(torch.from_numpy(w).float())
        self.w.bias = nn.Parameter(torch.zeros(time_dim))

    def forward(self, x: Tensor) -> Tensor:
        """Integer time deltas (the hot-path case: int64 differences of timestamps) go through
        `tgm_time2vec`; used standalone under autograd the weights receive their gradient
        (d/dw = -sin(w t + b) t, d/db = -sin(w t + b)).  Floating-point inputs, which the kernel's
        int64 interface cannot carry exactly, are evaluated with the same formula by device torch
        ops (no host sync to inspect the values)."""
        dev = _need_cuda(self.w.weight, 'Time2Vec')
        dt = x.to(device=dev)
        if dt.is_floating_point():
            return torch.cos(self.w(dt.float().unsqueeze(-1)))
        dt = dt.to(torch.int64).contiguous()
        if torch.is_grad_enabled() and (self.w.weight.requires_grad or self.w.bias.requires_grad):
            return _Time2VecFn.apply(dt, self.w.weight, self.w.bias, self.time_dim)
        with torch.no_grad():
            return _time2vec_kernel(dt, self.w.weight, self.w.bias, self.time_dim, dev)


def _time2vec_kernel(dt: Tensor, weight: Tensor, bias: Tensor, time_dim: int, dev) -> Tensor:
    out = torch.empty((*dt.shape, time_dim), dtype=torch.float32, device=dev)
    w = _f32(weight).reshape(-1)
    b = _f32(bias)
    _cabi.check(_cabi.lib.tgm_time2vec(dt.data_ptr(), dt.numel(), w.data_ptr(), b.data_ptr(),
                                       time_dim, out.data_ptr(), _cabi.current_stream(dev)))
    return out


class _Time2VecFn(torch.autograd.Function):
    """cos(w * float(dt) + b): forward on the CUDA kernel, backward -sin(arg) * (dt, 1)."""

    @staticmethod
    def forward(ctx, dt, weight, bias, time_dim):
        ctx.save_for_backward(dt, weight, bias)
        return _time2vec_kernel(dt, weight, bias, time_dim, dt.device)

    @staticmethod
    def backward(ctx, g):
        dt, weight, bias = ctx.saved_tensors
        tf = dt.float().reshape(-1, 1)
        s = -torch.sin(tf * weight.reshape(1, -1) + bias) * g.reshape(-1, weight.shape[0])
        return None, (s * tf).sum(0).reshape(weight.shape), s.sum(0), None


_DROPOUT_WARNED = set()


def warn_dropout_disabled(name: str, p: float) -> None:
    """The fused kernels of this package do not apply dropout (a mask drawn here could not follow
    the reference's RNG stream, and the backward passes are written for the dropout-free graph):
    a module constructed with the reference's default p > 0 trains with dropout DISABLED and says
    so once, instead of refusing the shipped examples' default configuration."""
    if name not in _DROPOUT_WARNED:
        _DROPOUT_WARNED.add(name)
        import warnings
        warnings.warn(f'{name}: dropout p={p} is not applied on the B200 path; training proceeds '
                      'with dropout disabled (set dropout=0 to silence this)', UserWarning,
                      stacklevel=3)


class TemporalAttention(nn.Module):
    """Multi-head temporal attention over a seed's sampled neighbours (one query per seed)."""

    def __init__(self, n_heads: int, node_dim: int, edge_dim: int, time_dim: int,
                 dropout: float = 0.1) -> None:
        super().__init__()
        if any(x <= 0 for x in [n_heads, node_dim, edge_dim, time_dim]):
            raise ValueError('n_heads,node_dim,edge_dim,time_dim,out_dim must be > 0')
        out_dim = node_dim + time_dim
        self.pad_dim = 0
        if out_dim % n_heads != 0:
            self.pad_dim = n_heads - out_dim % n_heads
            out_dim += self.pad_dim
        self.n_heads, self.head_dim, self.out_dim = n_heads, out_dim // n_heads, out_dim
        self.node_dim, self.edge_dim, self.time_dim = node_dim, edge_dim, time_dim
        key_dim = node_dim + edge_dim + time_dim
        self.W_Q = nn.Linear(out_dim, out_dim, bias=False)
        self.W_KV = nn.Linear(key_dim, out_dim * 2, bias=False)
        self.W_O = nn.Linear(out_dim, out_dim)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(out_dim)
        self._native = _NativeHandle(_cabi.lib.tgm_attn_destroy)

    def _handle(self, time_encoder: Time2Vec, dev: torch.device) -> ctypes.c_void_p:
        params = [self.W_Q.weight, self.W_KV.weight, self.W_O.weight, self.W_O.bias,
                  self.layer_norm.weight, self.layer_norm.bias, time_encoder.w.weight,
                  time_encoder.w.bias]
        ver = _version(params)
        if self._native.version != ver and self._native.h.value and \
                getattr(self, '_native_dev', None) == dev:
            # same shapes, new values (optimizer step): refresh the copies in place
            t = [_f32(p) for p in params]
            _cabi.check(_cabi.lib.tgm_attn_set_params(
                self._native.h, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(),
                t[4].data_ptr(), t[5].data_ptr(), t[6].reshape(-1).data_ptr(), t[7].data_ptr(),
                _cabi.current_stream(dev)))
            self._native.version = ver
        elif self._native.version != ver:
            self._native.free()
            for p in params:
                _need_cuda(p, 'TemporalAttention parameters')
            t = [_f32(p) for p in params]
            self._native_dev = dev
            _cabi.check(_cabi.lib.tgm_attn_create(
                ctypes.byref(self._native.h), self.n_heads, self.node_dim, self.edge_dim,
                self.time_dim, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(),
                t[4].data_ptr(), t[5].data_ptr(), float(self.layer_norm.eps),
                t[6].reshape(-1).data_ptr(), t[7].data_ptr(), dev.index))
            self._native.version = ver
        return self._native.h

    @torch.no_grad()
    def forward(self, node_x: Tensor, time_feat: Tensor, edge_feat: Tensor, nbr_node_feat: Tensor,
                nbr_time_feat: Tensor, valid_nbr_mask: Tensor,
                time_encoder: Optional[Time2Vec] = None) -> Tensor:
        """The reference signature (attention.py:58-66): time features supplied by the caller."""
        if self.training and self.dropout.p > 0:
            warn_dropout_disabled('TemporalAttention', self.dropout.p)
        dev = _need_cuda(node_x, 'TemporalAttention')
        S, k = valid_nbr_mask.shape
        te = time_encoder if time_encoder is not None else self._dummy_encoder(dev)
        nid = torch.where(valid_nbr_mask, 0, -1).to(torch.int32).contiguous()
        out = torch.empty((S, self.out_dim), dtype=torch.float32, device=dev)
        args = [_f32(node_x), _f32(time_feat), _f32(edge_feat), _f32(nbr_node_feat),
                _f32(nbr_time_feat), nid]
        _cabi.check(_cabi.lib.tgm_attn_forward_feats(
            self._handle(te, dev), *[a.data_ptr() for a in args], S, k, out.data_ptr(),
            _cabi.current_stream(dev)))
        return out

    def _dummy_encoder(self, dev: torch.device) -> Time2Vec:
        if getattr(self, '_te', None) is None:
            object.__setattr__(self, '_te', Time2Vec(self.time_dim).to(dev))
        return self._te

    def forward_fused(self, time_encoder: Time2Vec, node_x: Tensor, nbr_node_feat: Tensor,
                      edge_feat: Tensor, seed_times: Tensor, nbr_times: Tensor,
                      nbr_nids: Tensor) -> Tensor:
        """attention.py:58-128 with the Time2Vec calls of tgat.py:141-146 folded in: the time
        features are computed inside the kernel from (seed_times, nbr_times).  Differentiable
        when autograd is recording (dropout must be 0 in training mode)."""
        if self.training and self.dropout.p > 0:
            warn_dropout_disabled('TemporalAttention', self.dropout.p)
        dev = _need_cuda(node_x, 'TemporalAttention')
        params = [self.W_Q.weight, self.W_KV.weight, self.W_O.weight, self.W_O.bias,
                  self.layer_norm.weight, self.layer_norm.bias, time_encoder.w.weight,
                  time_encoder.w.bias]
        from tgm_b200.sampler import LazyEdgeRows
        lazy = edge_feat if isinstance(edge_feat, LazyEdgeRows) else None
        if torch.is_grad_enabled() and any(
                t.requires_grad for t in (*params, node_x, nbr_node_feat) +
                (() if lazy is not None else (edge_feat,))):
            if lazy is not None:
                edge_feat = lazy.materialize()  # the backward pass takes the feature block
            return _FusedAttention.apply(self, time_encoder, seed_times, nbr_times, nbr_nids,
                                         node_x, nbr_node_feat, edge_feat, *params)
        if lazy is not None:  # sampled rows read in place from the store's table, by edge id
            with torch.no_grad():
                S, k = nbr_nids.shape
                out = torch.empty((S, self.out_dim), dtype=torch.float32, device=dev)
                args = [_f32(node_x), _f32(nbr_node_feat), _f32(lazy.table),
                        lazy.rows.to(torch.int32).contiguous(),
                        seed_times.to(torch.int64).contiguous(),
                        nbr_times.to(torch.int64).contiguous(),
                        nbr_nids.to(torch.int32).contiguous()]
                _cabi.check(_cabi.lib.tgm_attn_forward_rows(
                    self._handle(time_encoder, dev), *[a.data_ptr() for a in args], S, k,
                    out.data_ptr(), _cabi.current_stream(dev)))
                return out
        with torch.no_grad():
            return self._forward_fused_nograd(time_encoder, node_x, nbr_node_feat, edge_feat,
                                              seed_times, nbr_times, nbr_nids, dev)

    def covers_hops(self, time_encoder: Time2Vec, k: int, dev: torch.device) -> bool:
        """True when `forward_hops` can serve this module (tgm_attn_folded_covers)."""
        return _cabi.lib.tgm_attn_folded_covers(self._handle(time_encoder, dev), int(k)) == 1

    def forward_hops(self, time_encoder: Time2Vec, node_x: Tensor, nbr_node_feat: Tensor,
                     edge_feats, seed_times: Tensor, nbr_times: Tensor,
                     nbr_nids: Tensor) -> Tensor:
        """Several hops of one TGAT layer in ONE call (no gradient): all arguments cover the hops'
        seeds back to back except the edge features, given per hop -- a list of dense
        (rows_i, k, edge_dim) blocks left where the sampler wrote them, or of LazyEdgeRows (their
        row ids are concatenated, the table is shared).  tgm_attn_forward_segments / _rows."""
        from tgm_b200.sampler import LazyEdgeRows
        dev = _need_cuda(node_x, 'TemporalAttention')
        S, k = nbr_nids.shape
        out = torch.empty((S, self.out_dim), dtype=torch.float32, device=dev)
        h = self._handle(time_encoder, dev)
        common = [seed_times.to(torch.int64).contiguous(), nbr_times.to(torch.int64).contiguous(),
                  nbr_nids.to(torch.int32).contiguous()]
        x, nf = _f32(node_x), _f32(nbr_node_feat)
        if all(isinstance(e, LazyEdgeRows) for e in edge_feats):
            table = _f32(edge_feats[0].table)
            rows = torch.cat([e.rows.to(torch.int32).reshape(-1, k) for e in edge_feats])
            _cabi.check(_cabi.lib.tgm_attn_forward_rows(
                h, x.data_ptr(), nf.data_ptr(), table.data_ptr(), rows.data_ptr(),
                *[a.data_ptr() for a in common], S, k, out.data_ptr(), _cabi.current_stream(dev)))
            return out
        segs = [_f32(e.materialize() if isinstance(e, LazyEdgeRows) else e) for e in edge_feats]
        n = len(segs)
        ptrs = (ctypes.c_void_p * n)(*[e.data_ptr() for e in segs])
        rows = (ctypes.c_int64 * n)(*[e.shape[0] for e in segs])
        _cabi.check(_cabi.lib.tgm_attn_forward_segments(
            h, x.data_ptr(), nf.data_ptr(), ptrs, rows, n, *[a.data_ptr() for a in common], S, k,
            out.data_ptr(), _cabi.current_stream(dev)))
        return out

    def _forward_fused_nograd(self, time_encoder, node_x, nbr_node_feat, edge_feat, seed_times,
                              nbr_times, nbr_nids, dev) -> Tensor:
        S, k = nbr_nids.shape
        out = torch.empty((S, self.out_dim), dtype=torch.float32, device=dev)
        args = [_f32(node_x), _f32(nbr_node_feat), _f32(edge_feat),
                seed_times.to(torch.int64).contiguous(), nbr_times.to(torch.int64).contiguous(),
                nbr_nids.to(torch.int32).contiguous()]
        _cabi.check(_cabi.lib.tgm_attn_forward(
            self._handle(time_encoder, dev), *[a.data_ptr() for a in args], S, k, out.data_ptr(),
            _cabi.current_stream(dev)))
        return out


class MergeLayer(nn.Module):
    """fc2(relu(fc1(cat[x1, x2])))."""

    def __init__(self, in_dim1: int, in_dim2: int, hidden_dim: int, output_dim: int) -> None:
        super().__init__()
        self.in_dim1, self.in_dim2 = in_dim1, in_dim2
        self.fc1 = nn.Linear(in_dim1 + in_dim2, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, output_dim)
        self._native = _NativeHandle(_cabi.lib.tgm_mlp2_destroy)

    def forward(self, x1: Tensor, x2: Tensor) -> Tensor:
        dev = _need_cuda(x1, 'MergeLayer')
        params = [self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias]
        if torch.is_grad_enabled() and any(t.requires_grad for t in (*params, x1, x2)):
            # two plain Linear layers: when autograd is recording they run as cuBLAS GEMMs
            # through torch (plain library GEMMs either way), which supplies their backward
            return self.fc2(self.fc1(torch.cat([x1, x2], dim=1)).relu())
        with torch.no_grad():
            return self._forward_nograd(x1, x2, dev, params)

    def _handle(self, dev: torch.device) -> ctypes.c_void_p:
        """The C handle for the current parameters (rebuilt when they change)."""
        params = [self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias]
        ver = _version(params)
        if self._native.version != ver:
            self._native.free()
            t = [_f32(p) for p in params]
            for p in t:
                _need_cuda(p, 'MergeLayer parameters')
            _cabi.check(_cabi.lib.tgm_mlp2_create(
                ctypes.byref(self._native.h), self.in_dim1, self.in_dim2, self.fc1.out_features,
                self.fc2.out_features, *[p.data_ptr() for p in t], dev.index))
            self._native.version = ver
        return self._native.h

    def _forward_nograd(self, x1: Tensor, x2: Tensor, dev, params) -> Tensor:
        h = self._handle(dev)
        S = x1.shape[0]
        out = torch.empty((S, self.fc2.out_features), dtype=torch.float32, device=dev)
        a, b = _f32(x1), _f32(x2)
        _cabi.check(_cabi.lib.tgm_mlp2_forward(h, a.data_ptr(), b.data_ptr(), S,
                                               out.data_ptr(), _cabi.current_stream(dev)))
        return out


def gather_rows(table: Tensor, ids: Tensor) -> Tensor:
    """table[ids] with torch's negative indexing (tgat.py:131-134), via tgm_gather_rows."""
    dev = _need_cuda(table, 'gather_rows')
    if torch.is_grad_enabled() and table.requires_grad:  # learnable node features: let autograd
        return table[ids.to(device=dev, dtype=torch.int64)]  # scatter the gradient rows
    table = _f32(table)
    ids = ids.to(device=dev, dtype=torch.int32).contiguous()
    out = torch.empty((ids.numel(), table.shape[1]), dtype=torch.float32, device=dev)
    _cabi.check(_cabi.lib.tgm_gather_rows(table.data_ptr(), table.shape[0], table.shape[1],
                                          ids.data_ptr(), ids.numel(), out.data_ptr(),
                                          _cabi.current_stream(dev)))
    return out


def masked_mean(z: Tensor, nbr_nids: Tensor) -> Tensor:
    """sum_k(z * mask) / clamp(sum mask, 1) over the sampled neighbours
    (examples/linkproppred/graphmixer.py:131-135), via tgm_masked_mean."""
    dev = _need_cuda(z, 'masked_mean')
    S, k, D = z.shape
    z = _f32(z)
    nid = nbr_nids.to(device=dev, dtype=torch.int32).contiguous()
    out = torch.empty((S, D), dtype=torch.float32, device=dev)  # every element is written
    _cabi.check(_cabi.lib.tgm_masked_mean(z.data_ptr(), nid.data_ptr(), S, k, D, out.data_ptr(),
                                          _cabi.current_stream(dev)))
    return out
