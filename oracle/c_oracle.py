"""ctypes face of the plain-C CPU oracle (oracle/recency_ring.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by tgm_b200/.

`CRing` restates RecencyNeighborHook's ring state machine (reference tgm-team/tgm @ 5183dc9,
tgm/hooks/neighbors/recency.py:93-97, 111-117, 239-321, 323-399); `run_stream` drives it the way
DGDataLoader + the hook do (recency.py:119-171, tgm/data/loader.py:136-160) and returns
position-sensitive checksums for full-size parity checks.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_int, c_int32, c_int64, c_uint64, c_void_p
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_build', 'librecency_oracle.so')
GOLD = 0x9E3779B97F4A7C15  # checksum weight multiplier (recency_ring.c)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, 'recency_ring.c')
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        proc = subprocess.run(['make', '-C', _HERE] + (['-B'] if force else []),
                              capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f'oracle build failed:\n{proc.stdout}\n{proc.stderr}')
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.ring_create.restype = c_void_p
        L.ring_create.argtypes = [c_int32, c_int32, c_int32]
        L.ring_destroy.argtypes = [c_void_p]
        L.ring_reset.argtypes = [c_void_p]
        L.ring_state.argtypes = [c_void_p] + [POINTER(c_void_p)] * 4
        L.ring_query.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p,
                                 c_void_p, c_void_p]
        L.ring_update.restype = c_int
        L.ring_update.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int]
        L.ring_run_stream.restype = c_int64
        L.ring_run_stream.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                      c_int64, c_int64, c_void_p, c_int32, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p]
        L.masked_mean_ref.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p]
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


class CRing:
    """Same surface as oracle.recency_oracle.RingOracle, C speed."""

    def __init__(self, num_nodes: int, num_nbrs: Sequence[int], edge_x_dim: int = 0,
                 directed: bool = False) -> None:
        if not len(num_nbrs):
            raise ValueError('num_nbrs must be non-empty')
        if not all(isinstance(x, int) and x > 0 for x in num_nbrs):
            raise ValueError('Each value in num_nbrs must be a positive integer')
        self.N, self.num_nbrs, self.B = int(num_nodes), list(num_nbrs), max(num_nbrs)
        self.D, self.directed = int(edge_x_dim), bool(directed)
        self._h = lib().ring_create(self.N, self.B, self.D)
        if not self._h:
            raise MemoryError('ring_create failed')

    def __del__(self):
        if getattr(self, '_h', None):
            lib().ring_destroy(self._h)
            self._h = None

    def reset_state(self) -> None:
        lib().ring_reset(self._h)

    def _state(self):
        p = [c_void_p() for _ in range(4)]
        lib().ring_state(self._h, *[ctypes.byref(q) for q in p])
        N, B, D = self.N, self.B, self.D

        def arr(q, ctype, shape):
            n = int(np.prod(shape))
            if n == 0:
                return np.zeros(shape, np.dtype(ctype))
            return np.ctypeslib.as_array(ctypes.cast(q, POINTER(ctype)), (n,)).reshape(shape).copy()
        return (arr(p[0], ctypes.c_int32, (N, B)), arr(p[1], ctypes.c_int64, (N, B)),
                arr(p[2], ctypes.c_float, (N, B, D)), arr(p[3], ctypes.c_int32, (N,)))

    ids = property(lambda s: s._state()[0])
    times = property(lambda s: s._state()[1])
    feats = property(lambda s: s._state()[2])
    write_pos = property(lambda s: s._state()[3])

    def query(self, seeds, tq, k: int):
        seeds, tq = _c(seeds, np.int32), _c(tq, np.int64)
        S = len(seeds)
        nid = np.empty((S, k), np.int32)
        nt = np.empty((S, k), np.int64)
        nx = np.empty((S, k, self.D), np.float32)
        lib().ring_query(self._h, _p(seeds), _p(tq), S, k, _p(nid), _p(nt), _p(nx))
        return nid, nt, nx

    def update(self, src, dst, t, x) -> None:
        src, dst, t = _c(src, np.int32), _c(dst, np.int32), _c(t, np.int64)
        x = None if (x is None or self.D == 0) else _c(x, np.float32)
        if lib().ring_update(self._h, _p(src), _p(dst), _p(t), _p(x), len(src), int(self.directed)):
            raise MemoryError('ring_update failed')

    def hook_call(self, seeds, tq, src, dst, t, x):
        """One RecencyNeighborHook.__call__ (recency.py:119-171)."""
        out = []
        if len(seeds):
            s, q = _c(seeds, np.int32), _c(tq, np.int64)
            for hop, k in enumerate(self.num_nbrs):
                if hop > 0:
                    s, q = out[-1][2].reshape(-1), out[-1][3].reshape(-1)
                nid, nt, nx = self.query(s, q, k)
                out.append((s, q, nid, nt, nx))
            if len(src):
                self.update(src, dst, t, x)
        return out

    def push_stream(self, src, dst, t, x, e_lo: int, e_hi: int, bs: int) -> None:
        """Push edges [e_lo, e_hi) batch by batch without querying (recency.py:323-399 only):
        the state a hook reaches after those batches, for timing runs that start mid-stream."""
        for lo in range(e_lo, e_hi, bs):
            hi = min(lo + bs, e_hi)
            self.update(src[lo:hi], dst[lo:hi], t[lo:hi], None if x is None else x[lo:hi])

    def run_stream(self, src, dst, t, x, e_lo: int, e_hi: int, bs: int, keep_hop0: bool = False,
                   checksum: bool = True):
        """Loader + hook over edges [e_lo, e_hi) with seeds [src | dst].  Returns
        (sampled slots, csum uint64[nhops, 3], hop-0 outputs or None); `checksum=False` skips the
        checksum pass over the outputs (timing runs)."""
        src, dst, t = _c(src, np.int32), _c(dst, np.int32), _c(t, np.int64)
        x = None if (x is None or self.D == 0) else _c(x, np.float32)
        nn = np.asarray(self.num_nbrs, np.int32)
        csum = np.zeros((len(nn), 3), np.uint64)
        outs = None
        if keep_hop0:
            S, k = 2 * (e_hi - e_lo), int(nn[0])
            outs = (np.empty((S, k), np.int32), np.empty((S, k), np.int64),
                    np.empty((S, k, self.D), np.float32))
        slots = lib().ring_run_stream(
            self._h, _p(src), _p(dst), _p(t), _p(x), e_lo, e_hi, bs, _p(nn), len(nn),
            int(self.directed), _p(csum) if checksum else None,
            *(map(_p, outs) if outs else (None, None, None)))
        if slots < 0:
            raise MemoryError('ring_run_stream failed')
        return int(slots), csum, outs


def masked_mean(z: np.ndarray, nid: np.ndarray) -> np.ndarray:
    z, nid = _c(z, np.float32), _c(nid, np.int32)
    S, k, D = z.shape
    out = np.empty((S, D), np.float32)
    lib().masked_mean_ref(_p(z), _p(nid), S, k, D, _p(out))
    return out


def checksum_np(v: np.ndarray, base: int = 0) -> int:
    """The checksum of recency_ring.c for an int32/int64/float32 block (numpy, wrapping)."""
    flat = v.reshape(-1)
    if flat.dtype == np.float32:
        flat = flat.view(np.int32)
    with np.errstate(over='ignore'):
        w = (np.arange(base, base + flat.size, dtype=np.uint64) * np.uint64(GOLD) + np.uint64(1))
        return int((flat.astype(np.int64).view(np.uint64) * w).sum(dtype=np.uint64))
