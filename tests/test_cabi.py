"""The C-ABI library loads on a CPU-only host, exports every symbol include/tgm_b200.h declares,
and its device entry points fail loudly (no CPU fallback) when no GPU is visible."""
import ctypes
import os
import re

import numpy as np
import pytest

from tgm_b200 import _cabi

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include',
                      'tgm_b200.h')


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(tgm_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_what_we_bind():
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert set(declared) == set(_cabi.SIGNATURES), set(declared) ^ set(_cabi.SIGNATURES)


@pytest.mark.parametrize('name', _declared_symbols())
def test_symbol_exported(name):
    assert getattr(_cabi.lib, name) is not None


def test_version_and_device_count():
    assert _cabi.lib.tgm_version() == 100
    assert _cabi.device_count() >= 0


def test_metadata_store_bounds_match_binary_search_semantics():
    """tgm_store_bounds == DGStorageArrayBackend._binary_search (array_backend.py:301-321)."""
    t = np.array([0, 0, 1, 1, 2, 5, 5, 5, 7, 9], np.int64)
    src = np.arange(10, dtype=np.int32)
    dst = src[::-1].copy()
    h = ctypes.c_void_p()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _cabi.check(_cabi.lib.tgm_store_create(ctypes.byref(h), P(src), P(dst), P(t), None, 10, 0, 10,
                                           -1, _cabi.TGM_MEM_HOST, None))

    def bounds(t_lo, t_hi, i_lo, i_hi):
        lb, ub = ctypes.c_int64(), ctypes.c_int64()
        _cabi.check(_cabi.lib.tgm_store_bounds(
            h, t_lo or 0, int(t_lo is not None), t_hi or 0, int(t_hi is not None),
            -1 if not i_lo else i_lo, -1 if not i_hi else i_hi, ctypes.byref(lb), ctypes.byref(ub)))
        return lb.value, ub.value

    def expect(t_lo, t_hi, i_lo, i_hi):
        lo = 0 if t_lo is None else int(np.searchsorted(t, t_lo, 'left'))
        hi = len(t) if t_hi is None else int(np.searchsorted(t, t_hi, 'right'))
        cl, ch = i_lo or 0, i_hi or len(t)
        return max(cl, min(ch, lo)), max(cl, min(ch, hi))

    for case in [(None, None, None, None), (1, 5, None, None), (1, 4, None, None),
                 (6, 6, None, None), (None, None, 2, 7), (1, 5, 4, 6), (0, 100, None, 3),
                 (10, None, None, None), (None, 0, None, None), (5, 5, 6, None)]:
        assert bounds(*case) == expect(*case), case
    # metadata-only store refuses to hand out device slabs
    rc = _cabi.lib.tgm_store_slab(h, 0, 1, None, None, None, None)
    assert rc == -3 and 'metadata-only' in _cabi.last_error()
    _cabi.lib.tgm_store_destroy(h)


def test_store_rejects_unsorted_and_bad_args():
    t = np.array([3, 1], np.int64)
    a = np.zeros(2, np.int32)
    h = ctypes.c_void_p()
    P = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    rc = _cabi.lib.tgm_store_create(ctypes.byref(h), P(a), P(a), P(t), None, 2, 0, 1, -1, 1, None)
    assert rc == -1 and 'non-decreasing' in _cabi.last_error()
    rc = _cabi.lib.tgm_store_create(ctypes.byref(h), P(a), P(a), P(t), None, 2, 0, 1, -1, 0, None)
    assert rc == -1  # a metadata-only store needs host arrays


def test_device_entry_points_fail_loudly_without_gpu():
    if _cabi.device_count() > 0:
        pytest.skip('a GPU is visible')
    h = ctypes.c_void_p()
    rc = _cabi.lib.tgm_recency_create(ctypes.byref(h), 8, 2, 0, 0)
    assert rc < 0 and not h.value
    with pytest.raises(_cabi.TGMNativeError):
        _cabi.require_device()


def test_new_entry_points_fail_loudly_without_gpu_or_arguments():
    """tgm_dedup_*, tgm_gae_*, tgm_recency_step: argument errors and the no-device case are
    reported through the status code + tgm_last_error, never by falling back to the host."""
    rc = _cabi.lib.tgm_recency_step(None, None, None, None, None, 1, 0, 1, None, None, None, None,
                                    None, None, None)
    assert rc == -1 and 'handle is NULL' in _cabi.last_error()
    rc = _cabi.lib.tgm_dedup_map(None, None, None, 8, None, 4, None, None)
    assert rc == -1 and 'NULL' in _cabi.last_error()
    assert _cabi.lib.tgm_set_option(b'gemm_fastf32', 1) == 0
    assert _cabi.lib.tgm_set_option(b'gemm_fastf32', 3) == -1
    assert _cabi.lib.tgm_set_option(b'no_such_option', 1) == -1
    if _cabi.device_count() > 0:
        return
    z = np.zeros(64, np.float32)
    P = z.ctypes.data_as(ctypes.c_void_p)
    h = ctypes.c_void_p()
    rc = _cabi.lib.tgm_gae_create(ctypes.byref(h), 4, 4, 2, 0, 2, *[P] * 11, 0)
    assert rc == -3 and not h.value and 'no such CUDA device' in _cabi.last_error()
    w, p, b = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    assert _cabi.lib.tgm_dedup_sizes(1000, ctypes.byref(w), ctypes.byref(p), ctypes.byref(b)) < 0


def test_tgn_training_entry_points_validate_their_arguments():
    """tgm_tgn_set_params / forward_saved / backward and tgm_gae_set_params / backward: a NULL
    handle is an argument error reported through the status code, with or without a GPU."""
    L = _cabi.lib
    calls = {
        'tgm_tgn_set_params': lambda: L.tgm_tgn_set_params(*[None] * 8),
        'tgm_tgn_forward_saved': lambda: L.tgm_tgn_forward_saved(None, None, 4, *[None] * 6),
        'tgm_tgn_backward': lambda: L.tgm_tgn_backward(None, None, None, None, 4, *[None] * 8),
        'tgm_tgn_set_aggregator': lambda: L.tgm_tgn_set_aggregator(None, 1, 0, None),
        'tgm_tgn_saved_aux_width': lambda: L.tgm_tgn_saved_aux_width(None),
        'tgm_gae_set_params': lambda: L.tgm_gae_set_params(*[None] * 13),
        'tgm_gae_backward': lambda: L.tgm_gae_backward(None, None, None, 4, None, None, None, None,
                                                       4, *[None] * 8)}
    for name, call in calls.items():
        assert call() == -1, name
        assert name in _cabi.last_error() and 'handle is NULL' in _cabi.last_error(), name
