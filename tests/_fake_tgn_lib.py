"""TEST INFRASTRUCTURE: a stand-in for the `tgm_tgn_*` / `tgm_gae_*` entry points of the C ABI,
implemented with the numpy oracle over HOST pointers, so that the host-side logic of
tgm_b200/nn/tgn.py (handle life cycle, parameter refresh, autograd routing, argument order, buffer
shapes) and the bodies of the GPU tests in tests/test_gpu_tgn_train.py can run on a CPU-only box.

It checks the Python face only; the CUDA kernels are checked by the `-m gpu` tests.  Nothing under
tgm_b200/ imports this.
"""
from __future__ import annotations

import contextlib
import ctypes
from typing import Dict

import numpy as np
import torch

from oracle import tgn_oracle as O

_I64, _I32, _F32 = (ctypes.c_int64, np.int64), (ctypes.c_int32, np.int32), (ctypes.c_float, np.float32)


def _arr(ptr, shape, kind=_F32):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, kind[1])
    assert ptr, 'NULL pointer for a non-empty array'
    return np.ctypeslib.as_array((kind[0] * n).from_address(int(ptr))).reshape(shape)


class _Tgn:
    def __init__(self, N, D, M, TD, ptrs):
        self.dims = (N, D, M, TD)
        self.orc = O.TGNMemoryOracle(N, D, M, TD, {})
        self.set_params(ptrs)

    def set_params(self, ptrs):
        N, D, M, TD = self.dims
        IN = D + 2 * M + TD
        shapes = [(3 * M, IN), (3 * M, M), (3 * M,), (3 * M,), (TD, 1), (TD,)]
        names = ['memory_updater.weight_ih', 'memory_updater.weight_hh', 'memory_updater.bias_ih',
                 'memory_updater.bias_hh', 'time_enc.w.weight', 'time_enc.w.bias']
        self.orc.p = {k: _arr(p, s).copy() for k, p, s in zip(names, ptrs, shapes)}


class _Gae:
    def __init__(self, IN, HC, H, D, TD, ptrs):
        self.dims = (IN, HC, H, D, TD)
        self.set_params(ptrs)

    def set_params(self, ptrs):
        IN, HC, H, D, TD = self.dims
        names = ['conv.lin_query.weight', 'conv.lin_query.bias', 'conv.lin_key.weight',
                 'conv.lin_key.bias', 'conv.lin_value.weight', 'conv.lin_value.bias',
                 'conv.lin_edge.weight', 'conv.lin_skip.weight', 'conv.lin_skip.bias',
                 'time_enc.w.weight', 'time_enc.w.bias']
        shapes = [(HC, IN), (HC,), (HC, IN), (HC,), (HC, IN), (HC,), (HC, TD + D), (HC, IN), (HC,),
                  (TD, 1), (TD,)]
        self.p = {k: _arr(p, s).copy() for k, p, s in zip(names, ptrs, shapes)}


class FakeLib:
    """Attribute access falls through to the real library for everything it does not fake."""

    def __init__(self, real):
        self._real = real
        self._objs: Dict[int, object] = {}
        self._next = 1000
        self.calls: Dict[str, int] = {}

    def __getattr__(self, name):
        return getattr(self._real, name)

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def _new(self, out_ref, obj):
        self._next += 1
        self._objs[self._next] = obj
        out_ref._obj.value = self._next
        return 0

    def _get(self, h):
        return self._objs[h.value if hasattr(h, 'value') else int(h)]

    # ---- TGN memory --------------------------------------------------------------------------
    def tgm_tgn_create(self, out, N, D, M, TD, *rest):
        self._count('tgm_tgn_create')
        return self._new(out, _Tgn(N, D, M, TD, rest[:6]))

    def tgm_tgn_destroy(self, h):
        self._objs.pop(h.value if hasattr(h, 'value') else int(h), None)

    def tgm_tgn_set_params(self, h, *rest):
        self._count('tgm_tgn_set_params')
        assert len(rest) == 7
        self._get(h).set_params(rest[:6])
        return 0

    def tgm_tgn_set_aggregator(self, h, kind, log_capacity, stream):
        self._count('tgm_tgn_set_aggregator')
        obj = self._get(h)
        obj.orc.aggregator = 'mean' if kind == 1 else 'last'
        obj.orc.reset_state()
        return 0

    def tgm_tgn_saved_aux_width(self, h):
        obj = self._get(h)
        return 2 * obj.dims[3] if obj.orc.aggregator == 'mean' else 2

    def _aux(self, obj, nid, n):
        """saved_aux as the library lays it out for the handle's aggregator."""
        N, D, M, TD = obj.dims
        aggr, rows, dt, weight = obj.orc.aggregated_messages(nid)
        if obj.orc.aggregator == 'mean':  # per row: mean of sin(arg) | mean of sin(arg) * dt
            p = obj.orc.p
            sn = np.sin(O._t2v_arg(dt, p['time_enc.w.weight'].reshape(-1), p['time_enc.w.bias']))
            x = np.asarray(dt).astype(np.float32).astype(np.float64)
            aux = np.zeros((n, 2 * TD))
            np.add.at(aux[:, :TD], rows, sn * weight[:, None])
            np.add.at(aux[:, TD:], rows, sn * (x * weight)[:, None])
            return aggr, aux
        aux = np.zeros((n, 2))
        aux[rows, 0], aux[rows, 1] = dt, 1.0
        return aggr, aux

    def tgm_tgn_reset(self, h, stream):
        self._get(h).orc.reset_state()
        return 0

    def tgm_tgn_state(self, h, pm, pl):
        orc = self._get(h).orc
        pm._obj.value = orc.memory.ctypes.data
        pl._obj.value = orc.last_update.ctypes.data
        return 0

    def tgm_tgn_forward(self, h, nid_p, n, training, mem_p, lu_p, stream):
        self._count('tgm_tgn_forward')
        if n == 0:  # as the library: nothing to do
            return 0
        obj = self._get(h)
        N, D, M, TD = obj.dims
        obj.orc.training = bool(training)
        z, lu = obj.orc.forward(_arr(nid_p, (n,), _I64))
        _arr(mem_p, (n, M))[:] = z
        _arr(lu_p, (n,), _I64)[:] = lu
        return 0

    def tgm_tgn_forward_saved(self, h, nid_p, n, mem_p, lu_p, sx_p, sh_p, aux_p, stream):
        self._count('tgm_tgn_forward_saved')
        obj = self._get(h)
        N, D, M, TD = obj.dims
        nid = _arr(nid_p, (n,), _I64)
        obj.orc.training = True
        z, lu = obj.orc.forward(nid)
        aggr, aux_rows = self._aux(obj, nid, n)
        _arr(mem_p, (n, M))[:] = z
        _arr(lu_p, (n,), _I64)[:] = lu
        _arr(sx_p, (n, D + 2 * M + TD))[:] = aggr
        _arr(sh_p, (n, M))[:] = obj.orc.memory[nid]
        _arr(aux_p, aux_rows.shape)[:] = aux_rows
        return 0

    def tgm_tgn_update_state(self, h, src_p, dst_p, t_p, raw_p, Eb, training, stream):
        obj = self._get(h)
        N, D, M, TD = obj.dims
        obj.orc.training = bool(training)
        obj.orc.update_state(_arr(src_p, (Eb,), _I32).astype(np.int64),
                             _arr(dst_p, (Eb,), _I32).astype(np.int64), _arr(t_p, (Eb,), _I64),
                             _arr(raw_p, (Eb, D)))
        return 0

    def tgm_tgn_flush(self, h, stream):
        orc = self._get(h).orc
        orc.training = True
        orc.train(False)
        return 0

    def tgm_tgn_backward(self, h, sx_p, sh_p, aux_p, n, dm_p, gwih, gwhh, gbih, gbhh, gtw, gtb,
                         stream):
        self._count('tgm_tgn_backward')
        obj = self._get(h)
        N, D, M, TD = obj.dims
        IN = D + 2 * M + TD
        p = obj.orc.p
        g = O.gru_cell_backward(p, _arr(sx_p, (n, IN)), _arr(sh_p, (n, M)), _arr(dm_p, (n, M)))
        d_enc = g.pop('x')[:, 2 * M + D:]
        if obj.orc.aggregator == 'mean':
            aux = _arr(aux_p, (n, 2 * TD)).astype(np.float64)
            gw, gb = -(aux[:, TD:] * d_enc).sum(0), -(aux[:, :TD] * d_enc).sum(0)
        else:
            aux = _arr(aux_p, (n, 2))
            gw, gb = O.time2vec_backward(aux[:, 0], p['time_enc.w.weight'].reshape(-1),
                                         p['time_enc.w.bias'], d_enc * aux[:, 1:2])
        _arr(gwih, (3 * M, IN))[:] += g['memory_updater.weight_ih']
        _arr(gwhh, (3 * M, M))[:] += g['memory_updater.weight_hh']
        _arr(gbih, (3 * M,))[:] += g['memory_updater.bias_ih']
        _arr(gbhh, (3 * M,))[:] += g['memory_updater.bias_hh']
        _arr(gtw, (TD,))[:] += gw.reshape(-1)
        _arr(gtb, (TD,))[:] += gb
        return 0

    # ---- TGN embedding -----------------------------------------------------------------------
    def tgm_gae_create(self, out, IN, HC, H, D, TD, *rest):
        self._count('tgm_gae_create')
        return self._new(out, _Gae(IN, HC, H, D, TD, rest[:11]))

    def tgm_gae_destroy(self, h):
        self._objs.pop(h.value if hasattr(h, 'value') else int(h), None)

    def tgm_gae_set_params(self, h, *rest):
        self._count('tgm_gae_set_params')
        assert len(rest) == 12
        self._get(h).set_params(rest[:11])
        return 0

    def _gae_inputs(self, obj, x_p, lu_p, n, es_p, ed_p, t_p, m_p, m):
        IN, HC, H, D, TD = obj.dims
        ei = np.stack([_arr(es_p, (m,), _I64), _arr(ed_p, (m,), _I64)])
        return (_arr(x_p, (n, IN)), _arr(lu_p, (n,), _I64), ei, _arr(t_p, (m,), _I64),
                _arr(m_p, (m, D)))

    def tgm_gae_forward(self, h, x_p, lu_p, n, es_p, ed_p, t_p, m_p, m, out_p, stream):
        self._count('tgm_gae_forward')
        obj = self._get(h)
        IN, HC, H, D, TD = obj.dims
        args = self._gae_inputs(obj, x_p, lu_p, n, es_p, ed_p, t_p, m_p, m)
        _arr(out_p, (n, HC))[:] = O.graph_attention_embedding(obj.p, H, *args)
        return 0

    def tgm_gae_backward(self, h, x_p, lu_p, n, es_p, ed_p, t_p, m_p, m, do_p, dx_p, gW, gb, gWe,
                         gtw, gtb, stream):
        self._count('tgm_gae_backward')
        obj = self._get(h)
        IN, HC, H, D, TD = obj.dims
        args = self._gae_inputs(obj, x_p, lu_p, n, es_p, ed_p, t_p, m_p, m)
        g = O.graph_attention_embedding_backward(obj.p, H, *args, _arr(do_p, (n, HC)))
        if dx_p:
            _arr(dx_p, (n, IN))[:] += g['x']
        W, b = _arr(gW, (4 * HC, IN)), _arr(gb, (4 * HC,))
        for q, nm in enumerate(('lin_query', 'lin_key', 'lin_value', 'lin_skip')):
            W[q * HC:(q + 1) * HC] += g[f'conv.{nm}.weight']
            b[q * HC:(q + 1) * HC] += g[f'conv.{nm}.bias']
        _arr(gWe, (HC, TD + D))[:] += g['conv.lin_edge.weight']
        _arr(gtw, (TD,))[:] += g['time_enc.w.weight'].reshape(-1)
        _arr(gtb, (TD,))[:] += g['time_enc.w.bias']
        return 0


@contextlib.contextmanager
def installed():
    """Route tgm_b200.nn.tgn through the fake library on CPU tensors; yields the FakeLib."""
    from tgm_b200 import _cabi
    from tgm_b200.nn import tgn

    def host_view(ptr_value, shape, dtype, device):
        kind = {torch.int32: _I32, torch.int64: _I64, torch.float32: _F32}[dtype]
        return torch.from_numpy(_arr(ptr_value, tuple(shape), kind))

    saved = (_cabi.lib, _cabi.current_stream, _cabi.device_view, tgn._require_cuda,
             tgn._device_index)
    fake = FakeLib(_cabi.lib)
    _cabi.lib, _cabi.current_stream, _cabi.device_view = fake, (lambda dev: 0), host_view
    tgn._require_cuda, tgn._device_index = (lambda dev, what: None), (lambda dev: 0)
    try:
        yield fake
    finally:
        (_cabi.lib, _cabi.current_stream, _cabi.device_view, tgn._require_cuda,
         tgn._device_index) = saved
