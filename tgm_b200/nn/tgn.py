"""TGNMemory on the B200 library: same constructor, parameter/buffer names and methods as
tgm/nn/encoder/tgn.py:80-251 (`forward`, `update_state`, `reset_state`, `detach`, `train`), with
IdentityMessage + LastAggregator.  The per-node Python dict message store and its per-node Python
loops (tgn.py:183-184, :226-229, :234) are a fixed-size device state behind `tgm_tgn_*`
(include/tgm_b200.h).  In training mode with autograd recording, `TGNMemory.forward` and
`GraphAttentionEmbedding.forward` are differentiable (`tgm_tgn_backward` / `tgm_gae_backward`
behind torch.autograd.Functions): the gradients the reference's training loop produces
(examples/linkproppred/tgn.py:100-118) for the GRU cell, the shared Time2Vec and the convolution.
The convolution's attention dropout must be 0 (it cannot follow the reference's RNG stream)."""
from __future__ import annotations

import ctypes
from typing import Callable, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from tgm_b200 import _cabi
from tgm_b200.nn.attention import Time2Vec, _NativeHandle, _f32, _version


def _require_cuda(dev: torch.device, what: str) -> None:
    if dev.type != 'cuda':
        raise _cabi.TGMNativeError(-3, f'{what} needs CUDA parameters (no CPU fallback)')


def _device_index(dev: torch.device) -> int:
    return dev.index if dev.index is not None else torch.cuda.current_device()


class IdentityMessage(nn.Module):
    def __init__(self, raw_msg_dim: int, memory_dim: int, time_dim: int) -> None:
        super().__init__()
        self.out_channels = raw_msg_dim + 2 * memory_dim + time_dim

    def forward(self, z_src: Tensor, z_dst: Tensor, raw_msg: Tensor, t_enc: Tensor) -> Tensor:
        return torch.cat([z_src, z_dst, raw_msg, t_enc], dim=-1)


class LastAggregator(nn.Module):
    """Marker module: the device state machine implements exactly this aggregator (tgn.py:43-56)."""


class MeanAggregator(nn.Module):
    """Marker module for scatter(mean) over a node's messages (tgn.py:59-63): selects
    TGM_TGN_AGGR_MEAN (an append-only device log of the batches since the last reset/flush)."""


class _TGNMemoryFn(torch.autograd.Function):
    """memory(n_id) in training mode as one differentiable op.  The reference calls
    loss.backward() after memory.update_state() (examples/linkproppred/tgn.py:111-118), so the
    rows the backward needs (GRU inputs, previous memory, time deltas) are saved at forward time;
    `tgm_tgn_backward` is a function of those rows and the current parameters only."""

    @staticmethod
    def forward(ctx, module, n_id, *params):
        dev = n_id.device
        n = n_id.numel()
        in_dim = module.raw_msg_dim + 2 * module.memory_dim + module.time_dim
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        mem, lu = f(n, module.memory_dim), torch.empty((n,), dtype=torch.int64, device=dev)
        handle = module._handle()
        aux_width = _cabi.lib.tgm_tgn_saved_aux_width(handle)  # 2 (last) | 2 * time_dim (mean)
        if aux_width <= 0:
            _cabi.check(aux_width)
        saved = (f(n, in_dim), f(n, module.memory_dim), f(n, aux_width))
        _cabi.check(_cabi.lib.tgm_tgn_forward_saved(
            handle, n_id.data_ptr(), n, mem.data_ptr(), lu.data_ptr(),
            *[t.data_ptr() for t in saved], _cabi.current_stream(dev)))
        ctx.module, ctx.saved_rows = module, saved
        ctx.mark_non_differentiable(lu)
        return mem, lu

    @staticmethod
    def backward(ctx, d_mem, _d_lu):
        module, saved = ctx.module, ctx.saved_rows
        dev = saved[0].device
        n, M, TD = saved[0].shape[0], module.memory_dim, module.time_dim
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        g = [z(3 * M, saved[0].shape[1]), z(3 * M, M), z(3 * M), z(3 * M), z(TD), z(TD)]
        d_mem = _f32(d_mem)
        _cabi.check(_cabi.lib.tgm_tgn_backward(
            module._handle(), *[t.data_ptr() for t in saved], n, d_mem.data_ptr(),
            *[t.data_ptr() for t in g], _cabi.current_stream(dev)))
        g[4] = g[4].reshape(TD, 1)  # Time2Vec weight is Linear(1, d).weight
        return (None, None, *g)


class TGNMemory(nn.Module):
    def __init__(self, num_nodes: int, raw_msg_dim: int, memory_dim: int, time_dim: int,
                 message_module: Optional[Callable] = None,
                 aggregator_module: Optional[Callable] = None) -> None:
        super().__init__()
        message_module = message_module or IdentityMessage(raw_msg_dim, memory_dim, time_dim)
        aggregator_module = aggregator_module or LastAggregator()
        if not isinstance(message_module, IdentityMessage) or \
                not isinstance(aggregator_module, (LastAggregator, MeanAggregator)):
            raise NotImplementedError('the B200 TGN memory implements IdentityMessage with the '
                                      'LastAggregator or the MeanAggregator (tgn.py:43-74)')
        self.num_nodes, self.raw_msg_dim = num_nodes, raw_msg_dim
        self.memory_dim, self.time_dim = memory_dim, time_dim
        self.msg_s_module, self.msg_d_module = message_module, message_module
        self.aggr_module = aggregator_module
        self.time_enc = Time2Vec(time_dim=time_dim)
        self.memory_updater = nn.GRUCell(message_module.out_channels, memory_dim)
        self._native = _NativeHandle(_cabi.lib.tgm_tgn_destroy)

    @property
    def device(self) -> torch.device:
        return self.time_enc.w.weight.device

    def _params(self):
        return [self.memory_updater.weight_ih, self.memory_updater.weight_hh,
                self.memory_updater.bias_ih, self.memory_updater.bias_hh,
                self.time_enc.w.weight, self.time_enc.w.bias]

    def _handle(self) -> ctypes.c_void_p:
        params = self._params()
        ver = _version(params)
        if self._native.version != ver and self._native.h.value and \
                getattr(self, '_native_dev', None) == self.device:
            # same shapes, new values (optimizer step): refresh the copies in place; the node
            # state and the message stores stay
            t = [_f32(p) for p in params]
            _cabi.check(_cabi.lib.tgm_tgn_set_params(
                self._native.h, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(),
                t[4].reshape(-1).data_ptr(), t[5].data_ptr(), _cabi.current_stream(self.device)))
            self._native.version = ver
        elif self._native.version != ver:
            dev = self.device
            _require_cuda(dev, 'TGNMemory')
            old = self._snapshot() if self._native.h.value else None
            self._native.free()
            t = [_f32(p) for p in params]
            _cabi.check(_cabi.lib.tgm_tgn_create(
                ctypes.byref(self._native.h), self.num_nodes, self.raw_msg_dim, self.memory_dim,
                self.time_dim, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(),
                t[4].reshape(-1).data_ptr(), t[5].data_ptr(),
                _device_index(dev)))
            if isinstance(self.aggr_module, MeanAggregator):
                _cabi.check(_cabi.lib.tgm_tgn_set_aggregator(self._native.h, 1, 0,
                                                             _cabi.current_stream(dev)))
            self._native.version = ver
            self._native_dev = dev
            pending = self.__dict__.pop('_pending_state', None)  # from load_state_dict
            if pending is not None:
                old = pending
            if old is not None:  # parameters moved to another device / a loaded checkpoint
                self.memory.copy_(old[0])
                self.last_update.copy_(old[1])
        return self._native.h

    def _views(self) -> Tuple[Tensor, Tensor]:
        pm, pl = ctypes.c_void_p(), ctypes.c_void_p()
        _cabi.check(_cabi.lib.tgm_tgn_state(self._handle(), ctypes.byref(pm), ctypes.byref(pl)))
        dev = self.device
        return (_cabi.device_view(pm.value, (self.num_nodes, self.memory_dim), torch.float32, dev),
                _cabi.device_view(pl.value, (self.num_nodes,), torch.int64, dev))

    def _snapshot(self):
        pm, pl = ctypes.c_void_p(), ctypes.c_void_p()
        _cabi.check(_cabi.lib.tgm_tgn_state(self._native.h, ctypes.byref(pm), ctypes.byref(pl)))
        dev = getattr(self, '_native_dev', None) or self.device  # where the handle lives
        return (_cabi.device_view(pm.value, (self.num_nodes, self.memory_dim), torch.float32,
                                  dev).clone(),
                _cabi.device_view(pl.value, (self.num_nodes,), torch.int64, dev).clone())

    # -- checkpoints: the reference registers `memory`, `last_update` and `_assoc` as buffers
    #    (tgn.py:128-133), so its state_dict carries them; here they live behind the handle
    def _save_to_state_dict(self, destination, prefix, keep_vars) -> None:
        super()._save_to_state_dict(destination, prefix, keep_vars)
        pending = self.__dict__.get('_pending_state')
        if self._native.h.value:
            mem, lu = self._snapshot()
        elif pending is not None:
            mem, lu = pending[0].clone(), pending[1].clone()
        else:  # no state yet: what reset_state() would give
            dev = self.device
            mem = torch.zeros((self.num_nodes, self.memory_dim), dtype=torch.float32, device=dev)
            lu = torch.zeros((self.num_nodes,), dtype=torch.int64, device=dev)
        destination[prefix + 'memory'] = mem
        destination[prefix + 'last_update'] = lu
        destination[prefix + '_assoc'] = torch.zeros((self.num_nodes,), dtype=torch.int64,
                                                     device=lu.device)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs) -> None:
        mem = state_dict.pop(prefix + 'memory', None)
        lu = state_dict.pop(prefix + 'last_update', None)
        state_dict.pop(prefix + '_assoc', None)  # scratch of _get_updated_memory upstream
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys,
                                      unexpected_keys, error_msgs)
        if mem is None and lu is None:
            return
        if mem is None or lu is None or tuple(mem.shape) != (self.num_nodes, self.memory_dim) or \
                tuple(lu.shape) != (self.num_nodes,):
            error_msgs.append(f'{prefix}memory / {prefix}last_update: expected shapes '
                              f'({self.num_nodes}, {self.memory_dim}) and ({self.num_nodes},)')
            return
        state = (mem.detach().to(torch.float32), lu.detach().to(torch.int64))
        if self._native.h.value:  # live handle: write through the views
            self.memory.copy_(state[0])
            self.last_update.copy_(state[1])
        else:  # applied when the handle is created (parameters may still be on the CPU)
            self.__dict__['_pending_state'] = state

    @property
    def memory(self) -> Tensor:
        """Live view of the device memory [num_nodes, memory_dim]."""
        return self._views()[0]

    @property
    def last_update(self) -> Tensor:
        return self._views()[1]

    def reset_state(self) -> None:
        _cabi.check(_cabi.lib.tgm_tgn_reset(self._handle(), _cabi.current_stream(self.device)))

    def detach(self) -> None:  # the device state never holds an autograd graph (tgn.py:154-155)
        return None

    def forward(self, n_id: Tensor) -> Tuple[Tensor, Tensor]:
        dev = self.device
        n_id = n_id.to(device=dev, dtype=torch.int64).contiguous()
        params = self._params()
        if self.training and n_id.numel() > 0 and torch.is_grad_enabled() and \
                any(p.requires_grad for p in params):
            return _TGNMemoryFn.apply(self, n_id, *params)
        with torch.no_grad():
            return self._forward_nograd(n_id, dev)

    def _forward_nograd(self, n_id: Tensor, dev: torch.device) -> Tuple[Tensor, Tensor]:
        n = n_id.numel()
        mem = torch.empty((n, self.memory_dim), dtype=torch.float32, device=dev)
        lu = torch.empty((n,), dtype=torch.int64, device=dev)
        _cabi.check(_cabi.lib.tgm_tgn_forward(self._handle(), n_id.data_ptr(), n,
                                              int(self.training), mem.data_ptr(), lu.data_ptr(),
                                              _cabi.current_stream(dev)))
        return mem, lu

    @torch.no_grad()
    def update_state(self, src: Tensor, dst: Tensor, t: Tensor, raw_msg: Tensor) -> None:
        dev = self.device
        src = src.to(device=dev, dtype=torch.int32).contiguous()
        dst = dst.to(device=dev, dtype=torch.int32).contiguous()
        t = t.to(device=dev, dtype=torch.int64).contiguous()
        raw = _f32(raw_msg.to(dev))
        _cabi.check(_cabi.lib.tgm_tgn_update_state(
            self._handle(), src.data_ptr(), dst.data_ptr(), t.data_ptr(), raw.data_ptr(),
            src.numel(), int(self.training), _cabi.current_stream(dev)))

    def train(self, mode: bool = True) -> 'TGNMemory':
        if self.training and not mode and self._native.h.value:
            # flush the message store into memory when entering eval mode (tgn.py:245-251)
            _cabi.check(_cabi.lib.tgm_tgn_flush(self._handle(), _cabi.current_stream(self.device)))
        return super().train(mode)


class TransformerConv(nn.Module):
    """Parameter container with torch_geometric.nn.TransformerConv's names and shapes (lin_key,
    lin_query, lin_value, lin_skip: Linear(in, heads*out) with bias; lin_edge: Linear(edge_dim,
    heads*out) without), so a PyG state_dict loads into it.  The arithmetic runs in
    `tgm_gae_forward` (GraphAttentionEmbedding below)."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, dropout: float = 0.0,
                 edge_dim: Optional[int] = None) -> None:
        super().__init__()
        if edge_dim is None:
            raise NotImplementedError('the B200 TransformerConv is the edge-attributed form '
                                      'GraphAttentionEmbedding uses (tgn.py:25-27)')
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.dropout, self.edge_dim = dropout, edge_dim
        self.lin_key = nn.Linear(in_channels, heads * out_channels)
        self.lin_query = nn.Linear(in_channels, heads * out_channels)
        self.lin_value = nn.Linear(in_channels, heads * out_channels)
        self.lin_edge = nn.Linear(edge_dim, heads * out_channels, bias=False)
        self.lin_skip = nn.Linear(in_channels, heads * out_channels)


class _GAEFn(torch.autograd.Function):
    """tgm_gae_forward / tgm_gae_backward as one differentiable op; the backward recomputes the
    forward from the inputs, so nothing else is saved."""

    @staticmethod
    def forward(ctx, module, x, lu, ei, tt, mm, *params):
        dev = x.device
        xs = _f32(x)
        out = module._run_forward(xs, lu, ei, tt, mm, dev)
        ctx.module, ctx.args, ctx.need_x = module, (xs, lu, ei, tt, mm), x.requires_grad
        return out

    @staticmethod
    def backward(ctx, d_out):
        module = ctx.module
        x, lu, ei, tt, mm = ctx.args
        dev = x.device
        c = module.conv
        HC, n, m = c.heads * c.out_channels, x.shape[0], ei.shape[1]
        TD, A = module.time_enc.time_dim, module.msg_dim + module.time_enc.time_dim
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        d_x = z(n, c.in_channels) if ctx.need_x else None
        gW, gb, gWe, gtw, gtb = z(4 * HC, c.in_channels), z(4 * HC), z(HC, A), z(TD), z(TD)
        d_out = _f32(d_out)
        _cabi.check(_cabi.lib.tgm_gae_backward(
            module._handle(dev), x.data_ptr(), lu.data_ptr(), n, ei[0].data_ptr(),
            ei[1].data_ptr(), tt.data_ptr(), mm.data_ptr(), m, d_out.data_ptr(), _cabi.ptr(d_x),
            gW.data_ptr(), gb.data_ptr(), gWe.data_ptr(), gtw.data_ptr(), gtb.data_ptr(),
            _cabi.current_stream(dev)))
        # parameter order of GraphAttentionEmbedding._params(); the stacked buffers hold the
        # query, key, value and skip linears in that order
        Wq, Wk, Wv, Ws = gW.split(HC)
        bq, bk, bv, bs = gb.split(HC)
        return (None, d_x, None, None, None, None,
                Wq, bq, Wk, bk, Wv, bv, gWe, Ws, bs, gtw.reshape(TD, 1), gtb)


class GraphAttentionEmbedding(nn.Module):
    """tgm/nn/encoder/tgn.py:14-40 on the B200 library: same constructor, `time_enc` shared with
    the memory module, `conv` carrying the TransformerConv parameters.  Differentiable in training
    mode when `conv.dropout == 0` (the convolution's attention dropout cannot follow the
    reference's RNG stream; with dropout > 0 training mode is refused).
    Parity with torch_geometric is unpinned: see include/tgm_b200.h (tgm_gae_*)."""

    def __init__(self, in_channels: int, out_channels: int, msg_dim: int, time_enc: nn.Module) -> None:
        super().__init__()
        self.time_enc = time_enc
        edge_dim = msg_dim + time_enc.time_dim
        self.msg_dim = msg_dim
        self.conv = TransformerConv(in_channels, out_channels // 2, heads=2, dropout=0.1,
                                    edge_dim=edge_dim)
        self._native = _NativeHandle(_cabi.lib.tgm_gae_destroy)

    def _params(self):
        c = self.conv
        return [c.lin_query.weight, c.lin_query.bias, c.lin_key.weight, c.lin_key.bias,
                c.lin_value.weight, c.lin_value.bias, c.lin_edge.weight, c.lin_skip.weight,
                c.lin_skip.bias, self.time_enc.w.weight, self.time_enc.w.bias]

    def _handle(self, dev: torch.device) -> ctypes.c_void_p:
        c = self.conv
        params = self._params()
        ver = _version(params)
        if self._native.version != ver and self._native.h.value and \
                getattr(self, '_native_dev', None) == dev:
            # same shapes, new values (optimizer step): refresh the copies in place
            t = [_f32(p) for p in params]
            t[9] = t[9].reshape(-1)
            _cabi.check(_cabi.lib.tgm_gae_set_params(self._native.h, *[p.data_ptr() for p in t],
                                                     _cabi.current_stream(dev)))
            self._native.version = ver
        elif self._native.version != ver:
            self._native.free()
            _require_cuda(dev, 'GraphAttentionEmbedding')
            t = [_f32(p) for p in params]
            t[9] = t[9].reshape(-1)
            _cabi.check(_cabi.lib.tgm_gae_create(
                ctypes.byref(self._native.h), c.in_channels, c.heads * c.out_channels, c.heads,
                self.msg_dim, self.time_enc.time_dim, *[p.data_ptr() for p in t],
                _device_index(dev)))
            self._native.version = ver
            self._native_dev = dev
        return self._native.h

    def _run_forward(self, x: Tensor, lu: Tensor, ei: Tensor, tt: Tensor, mm: Tensor,
                     dev: torch.device) -> Tensor:
        n, m = x.shape[0], ei.shape[1]
        out = torch.empty((n, self.conv.heads * self.conv.out_channels), dtype=torch.float32,
                          device=dev)
        _cabi.check(_cabi.lib.tgm_gae_forward(
            self._handle(dev), x.data_ptr(), lu.data_ptr(), n, ei[0].data_ptr(), ei[1].data_ptr(),
            tt.data_ptr(), mm.data_ptr(), m, out.data_ptr(), _cabi.current_stream(dev)))
        return out

    def forward(self, x: Tensor, last_update: Tensor, edge_index: Tensor, t: Tensor,
                msg: Tensor) -> Tensor:
        dev = self.conv.lin_key.weight.device
        if self.training and self.conv.dropout > 0:
            from tgm_b200.nn.attention import warn_dropout_disabled
            warn_dropout_disabled('GraphAttentionEmbedding', self.conv.dropout)
        x = x.to(dev)
        lu = last_update.to(device=dev, dtype=torch.int64).contiguous()
        ei = edge_index.to(device=dev, dtype=torch.int64).contiguous()
        tt = t.to(device=dev, dtype=torch.int64).contiguous()
        mm = _f32(msg.to(dev)).reshape(ei.shape[1], self.msg_dim)
        params = self._params()
        if self.training and x.shape[0] > 0 and torch.is_grad_enabled() and \
                (x.requires_grad or any(p.requires_grad for p in params)):
            return _GAEFn.apply(self, x, lu, ei, tt, mm, *params)
        with torch.no_grad():
            return self._run_forward(_f32(x), lu, ei, tt, mm, dev)
