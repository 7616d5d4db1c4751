"""ctypes binding of the C ABI declared in include/tgm_b200.h.

The product path has no CPU fallback: if the shared library is missing this module raises at
import, and every device entry point raises `TGMNativeError` when no CUDA device is present.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, byref, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64,
                    c_void_p)

import torch  # noqa: F401  loaded first so its bundled CUDA libraries (cuBLAS) are the ones resolved

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TGM_B200_LIB') or os.path.join(_HERE, 'csrc', 'libtgm_b200.so')

TGM_MEM_DEVICE, TGM_MEM_HOST = 0, 1


class TGMNativeError(RuntimeError):
    """A C-ABI call returned a negative status."""

    def __init__(self, code: int, message: str) -> None:
        super().__init__(f'[tgm_b200 rc={code}] {message}')
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f'{LIB_PATH} is missing. Build it with `python -m tgm_b200.build` (needs nvcc); '
        'tgm_b200 has no CPU fallback.'
    )

lib = ctypes.CDLL(LIB_PATH)

# every symbol include/tgm_b200.h declares, with its signature (checked by tests/test_cabi.py)
SIGNATURES = {
    'tgm_last_error': (c_char_p, []),
    'tgm_version': (c_int, []),
    'tgm_device_count': (c_int, []),
    'tgm_set_option': (c_int, [c_char_p, c_int]),
    'tgm_store_create': (c_int, [POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_int32, c_int32, c_int, c_int, c_void_p]),
    'tgm_store_destroy': (None, [c_void_p]),
    'tgm_store_info': (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int32), POINTER(c_int32),
                               POINTER(c_int)]),
    'tgm_store_bounds': (c_int, [c_void_p, c_int64, c_int, c_int64, c_int, c_int64, c_int64,
                                 POINTER(c_int64), POINTER(c_int64)]),
    'tgm_store_slab': (c_int, [c_void_p, c_int64, c_int64, POINTER(c_void_p), POINTER(c_void_p),
                               POINTER(c_void_p), POINTER(c_void_p)]),
    'tgm_recency_create': (c_int, [POINTER(c_void_p), c_int32, c_int32, c_int32, c_int]),
    'tgm_recency_destroy': (None, [c_void_p]),
    'tgm_recency_reset': (c_int, [c_void_p, c_void_p]),
    'tgm_recency_query': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    'tgm_recency_update': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                   c_int, c_void_p]),
    'tgm_recency_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                 c_int32, POINTER(c_int32), c_void_p, c_void_p, POINTER(c_void_p),
                                 POINTER(c_void_p), POINTER(c_void_p), c_void_p]),
    'tgm_recency_state': (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_void_p),
                                  POINTER(c_void_p), POINTER(c_void_p)]),
    'tgm_csr_build': (c_int, [POINTER(c_void_p), c_void_p, c_int64, c_int64, c_int, c_int,
                              c_void_p]),
    'tgm_csr_destroy': (None, [c_void_p]),
    'tgm_csr_info': (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64),
                             POINTER(c_int), POINTER(c_int)]),
    'tgm_csr_sample': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
                               c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    'tgm_csr_sample_edges': (c_int, [c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    'tgm_recency_dims': (c_int, [c_void_p, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    'tgm_csr_export_ring': (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    'tgm_csr_sample_uniform': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32,
                                       c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]),
    'tgm_csr_sample_uniform_time': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64,
                                            c_int, c_int32, c_uint64, c_void_p, c_void_p, c_void_p,
                                            c_void_p]),
    'tgm_csr_candidate_counts': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                         c_void_p]),
    'tgm_csr_gather_picks': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'tgm_csr_sample_edges_host': (c_int, [c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_int, c_void_p]),
    'tgm_csr_sample_ids': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
                                   c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    'tgm_csr_sample_edges_ids': (c_int, [c_void_p, c_int64, c_int64, c_int32, c_int32, c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p]),
    'tgm_csr_sample_edges_mean': (c_int, [c_void_p, c_int64, c_int64, c_int32, c_int32, c_int,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'tgm_csr_sample_edges_host_ids': (c_int, [c_void_p, c_int64, c_int64, c_int32, c_int32,
                                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_int, c_void_p]),
    'tgm_csr_sample_edges_host_mean': (c_int, [c_void_p, c_int64, c_int64, c_int32, c_int32,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_int, c_void_p]),
    'tgm_tc_linear': (c_int, [c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int, c_void_p, c_void_p]),
    'tgm_fastf32_linear': (c_int, [c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_void_p, c_void_p]),
    'tgm_join_row_bytes': (c_int64, [c_int32]),
    'tgm_join_pack': (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_int64, c_void_p,
                              c_void_p]),
    'tgm_join_scatter': (c_int, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                 c_void_p]),
    'tgm_negatives_window': (c_int, [c_uint64, c_uint64, c_int64, c_int64, c_int64, c_int64,
                                     c_void_p, c_void_p]),
    'tgm_frontier_compact': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    'tgm_masked_mean': (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p,
                                c_void_p]),
    'tgm_time2vec': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p,
                             c_void_p]),
    'tgm_attn_create': (c_int, [POINTER(c_void_p), c_int32, c_int32, c_int32, c_int32, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                c_void_p, c_void_p, c_int]),
    'tgm_attn_set_params': (c_int, [c_void_p] * 10),
    'tgm_attn_destroy': (None, [c_void_p]),
    'tgm_attn_out_dim': (c_int, [c_void_p]),
    'tgm_attn_forward': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    'tgm_attn_forward_rows': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    'tgm_attn_folded_covers': (c_int, [c_void_p, c_int32]),
    'tgm_attn_forward_segments': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                          c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p,
                                          c_void_p]),
    'tgm_small_gemm': (c_int, [c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32,
                               c_void_p, c_void_p]),
    'tgm_tgat_create': (c_int, [POINTER(c_void_p), c_int32, c_void_p, c_void_p, c_int]),
    'tgm_tgat_destroy': (None, [c_void_p]),
    'tgm_tgat_forward': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    'tgm_attn_forward_feats': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    'tgm_attn_backward': (c_int, [c_void_p] + [c_void_p] * 6 + [c_int64, c_int32] + [c_void_p] * 12 +
                          [c_void_p]),
    'tgm_mlp2_create': (c_int, [POINTER(c_void_p), c_int32, c_int32, c_int32, c_int32, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_int]),
    'tgm_mlp2_destroy': (None, [c_void_p]),
    'tgm_mlp2_forward': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'tgm_tgn_create': (c_int, [POINTER(c_void_p), c_int32, c_int32, c_int32, c_int32, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    'tgm_tgn_destroy': (None, [c_void_p]),
    'tgm_tgn_reset': (c_int, [c_void_p, c_void_p]),
    'tgm_tgn_state': (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_void_p)]),
    'tgm_tgn_forward': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                                c_void_p]),
    'tgm_tgn_update_state': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                     c_int, c_void_p]),
    'tgm_tgn_flush': (c_int, [c_void_p, c_void_p]),
    'tgm_tgn_set_params': (c_int, [c_void_p] * 8),
    'tgm_tgn_set_aggregator': (c_int, [c_void_p, c_int, c_int64, c_void_p]),
    'tgm_tgn_saved_aux_width': (c_int, [c_void_p]),
    'tgm_tgn_forward_saved': (c_int, [c_void_p, c_void_p, c_int64] + [c_void_p] * 6),
    'tgm_tgn_backward': (c_int, [c_void_p] * 4 + [c_int64] + [c_void_p] * 8),
    'tgm_dyg_create': (c_int, [POINTER(c_void_p), c_void_p, c_int]),
    'tgm_dyg_destroy': (None, [c_void_p]),
    'tgm_dyg_set_params': (c_int, [c_void_p, c_void_p, c_void_p]),
    'tgm_dyg_backward': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    'tgm_dyg_forward': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                c_void_p]),
    'tgm_gae_create': (c_int, [POINTER(c_void_p)] + [c_int32] * 5 + [c_void_p] * 11 + [c_int]),
    'tgm_gae_destroy': (None, [c_void_p]),
    'tgm_gae_forward': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'tgm_gae_set_params': (c_int, [c_void_p] * 13),
    'tgm_gae_backward': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_int64] + [c_void_p] * 8),
    'tgm_dedup_sizes': (c_int, [c_int32, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
    'tgm_dedup_unique': (c_int, [POINTER(c_void_p), POINTER(c_int64), POINTER(c_int32), c_int32,
                                 c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                 c_void_p, c_void_p]),
    'tgm_dedup_map': (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_void_p,
                              c_void_p]),
    'tgm_gather_rows': (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p,
                                c_void_p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


class DygLayer(ctypes.Structure):
    """tgm_dyg_layer (include/tgm_b200.h)."""
    _fields_ = [(n, c_void_p) for n in ('in_proj_w', 'in_proj_b', 'out_proj_w', 'out_proj_b',
                                        'ffn1_w', 'ffn1_b', 'ffn2_w', 'ffn2_b', 'ln0_w', 'ln0_b',
                                        'ln1_w', 'ln1_b')]


class DygParams(ctypes.Structure):
    """tgm_dyg_params (include/tgm_b200.h)."""
    _fields_ = ([(n, c_int32) for n in ('node_dim', 'edge_dim', 'time_dim', 'channel_dim',
                                        'out_dim', 'patch_size', 'num_layers', 'num_heads',
                                        'seq_len')] +
                [('ln_eps', c_float)] +
                [(n, c_void_p) for n in ('t2v_w', 't2v_b', 'cooc_w1', 'cooc_b1', 'cooc_w2',
                                         'cooc_b2')] +
                [('proj_w', c_void_p * 4), ('proj_b', c_void_p * 4),
                 ('layers', POINTER(DygLayer)), ('out_w', c_void_p), ('out_b', c_void_p)])


class DygGrads(ctypes.Structure):
    """tgm_dyg_grads (include/tgm_b200.h); the per-layer table has DygLayer's field order."""
    _fields_ = ([(n, c_void_p) for n in ('t2v_w', 't2v_b', 'cooc_w1', 'cooc_b1', 'cooc_w2',
                                         'cooc_b2')] +
                [('proj_w', c_void_p * 4), ('proj_b', c_void_p * 4),
                 ('layers', POINTER(DygLayer)), ('out_w', c_void_p), ('out_b', c_void_p)])


def last_error() -> str:
    msg = lib.tgm_last_error()
    return msg.decode('utf-8', 'replace') if msg else ''


def check(rc: int) -> None:
    if rc != 0:
        raise TGMNativeError(rc, last_error())


def device_count() -> int:
    return int(lib.tgm_device_count())


def require_device() -> None:
    if device_count() < 1:
        raise TGMNativeError(-3, 'no CUDA device visible: tgm_b200 has no CPU fallback')


def ptr(t) -> int | None:
    """data_ptr of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def current_stream(device) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


class _DevicePointer:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can alias it."""

    def __init__(self, ptr_value: int, shape, typestr: str) -> None:
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr,
                                         'data': (int(ptr_value), False), 'version': 2}


def device_view(ptr_value: int, shape, dtype, device):
    """Zero-copy torch view of library-owned device memory (valid while the handle lives)."""
    import torch
    typestr = {torch.int32: '<i4', torch.int64: '<i8', torch.float32: '<f4'}[dtype]
    return torch.as_tensor(_DevicePointer(ptr_value, shape, typestr), device=device)
