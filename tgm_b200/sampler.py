"""Stateless windowed recency sampler: the throughput form of the hot path.

`RecencyCSR` wraps `tgm_csr_*` (include/tgm_b200.h).  It answers exactly what
RecencyNeighborHook answers batch by batch (tgm/hooks/neighbors/recency.py:119-171 of the
reference), but one launch per hop serves every loader batch of a window of the event stream;
the batch a seed belongs to is encoded as an edge-index cut.  This is the path that is sharded
across GPUs by time range (tgm_b200/parallel.py) and that bench.py measures.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
from torch import Tensor

from tgm_b200 import _cabi


class LazyEdgeRows:
    """`nbr_edge_x` of a hop that was never materialised: `rows[s, c]` is the store
    edge whose feature row slot (s, c) holds (-1 = padding -> zeros), `table` the store's
    float32 [E, D] feature table.  The fused consumers (`TemporalAttention.forward_fused` ->
    tgm_attn_forward_rows) read the rows in place; anything else sees an ordinary tensor: torch
    functions and tensor attributes materialise it on first use (`materialize()`, cached)."""

    __slots__ = ('table', 'rows', '_dense')

    def __init__(self, table: Tensor, rows: Tensor) -> None:
        self.table, self.rows, self._dense = table, rows, None

    @property
    def shape(self) -> torch.Size:
        return torch.Size((*self.rows.shape, self.table.shape[1]))

    @property
    def dtype(self) -> torch.dtype:
        return self.table.dtype

    @property
    def device(self) -> torch.device:
        return self.rows.device

    @property
    def ndim(self) -> int:
        return self.rows.ndim + 1

    def size(self, dim: Optional[int] = None):
        return self.shape if dim is None else self.shape[dim]

    def __len__(self) -> int:
        return self.rows.shape[0]

    def materialize(self) -> Tensor:
        if self._dense is None:
            r = self.rows
            self._dense = self.table[r.clamp(min=0).long()] * (r >= 0).unsqueeze(-1)
        return self._dense

    def split(self, n: int):
        return tuple(LazyEdgeRows(self.table, r) for r in self.rows.split(n))

    def __getitem__(self, idx):
        if isinstance(idx, slice):  # a row range stays lazy
            return LazyEdgeRows(self.table, self.rows[idx])
        return self.materialize()[idx]

    def __getattr__(self, name):  # any other tensor attribute / method
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def dense(v):
            if isinstance(v, LazyEdgeRows):
                return v.materialize()
            if isinstance(v, (list, tuple)):
                return type(v)(dense(u) for u in v)
            return v
        return func(*dense(args), **{k: dense(v) for k, v in (kwargs or {}).items()})


def _delegate(name):
    def op(self, *args, **kwargs):
        return getattr(self.materialize(), name)(*args, **kwargs)
    op.__name__ = name
    return op


# operators are looked up on the type, not through __getattr__: forward them to the dense tensor
for _n in ('add', 'radd', 'sub', 'rsub', 'mul', 'rmul', 'truediv', 'rtruediv', 'matmul', 'rmatmul',
           'neg', 'pow', 'eq', 'ne', 'lt', 'le', 'gt', 'ge', 'iter', 'array', 'bool', 'float'):
    setattr(LazyEdgeRows, f'__{_n}__', _delegate(f'__{_n}__'))
LazyEdgeRows.__hash__ = object.__hash__


@dataclass
class HopSample:
    """One hop of sampled neighbourhoods (rows follow the reference's per-batch seed order)."""
    seed_nids: Tensor      # int32 (S,)
    seed_times: Tensor     # int64 (S,)
    nbr_nids: Tensor       # int32 (S, k)
    nbr_edge_time: Tensor  # int64 (S, k)
    nbr_edge_x: Tensor     # float32 (S, k, D), or LazyEdgeRows (sample_window(lazy_edge_x=True))


class RecencyCSR:
    """Per-node chronological adjacency of a device store for one loader geometry.

    Args:
        storage: a `DeviceCOOStorage` on a CUDA device.
        batch_size: loader batch size in events (batch_unit='r'); batches start at `e_start`.
        directed: only src->dst entries are pushed (RecencyNeighborHook(directed=True)).
        colocate_x: keep a copy of the feature rows in adjacency order so a seed's window is a
            contiguous read (2*E*D*4 bytes of HBM when undirected).
    """

    def __init__(self, storage, batch_size: int, directed: bool = False,
                 colocate_x: bool = True, e_start: int = 0) -> None:
        if storage.device is None:
            raise _cabi.TGMNativeError(-3, 'RecencyCSR needs a device-resident store')
        self._storage = storage  # keeps the store (and its slabs) alive
        self.device = storage.device
        self.batch_size = int(batch_size)
        self.directed = bool(directed)
        self.e_start = int(e_start)
        self.D = storage.get_edge_x_dim() or 0
        self._handle = ctypes.c_void_p()
        _cabi.check(_cabi.lib.tgm_csr_build(
            ctypes.byref(self._handle), storage.handle, self.e_start, self.batch_size,
            int(self.directed), int(colocate_x), _cabi.current_stream(self.device)))

    def __del__(self, _destroy=_cabi.lib.tgm_csr_destroy) -> None:
        h = getattr(self, '_handle', None)
        if h is not None and h.value:
            _destroy(h)
            h.value = None

    @property
    def handle(self) -> ctypes.c_void_p:
        return self._handle

    def _alloc(self, S: int, k: int):
        dev = self.device
        return (torch.empty((S, k), dtype=torch.int32, device=dev),
                torch.empty((S, k), dtype=torch.int64, device=dev),
                torch.empty((S, k, self.D), dtype=torch.float32, device=dev))

    def sample(self, seeds: Tensor, tq: Tensor, cut: Tensor, k: int, B: int,
               cut_group: int = 1, out=None):
        """General seeds: seed i sees entries of edges < cut[i // cut_group]."""
        S = seeds.numel()
        nid, nt, nx = out if out is not None else self._alloc(S, k)
        _cabi.check(_cabi.lib.tgm_csr_sample(
            self._handle, seeds.data_ptr(), tq.data_ptr(), cut.data_ptr(), int(cut_group), S,
            int(B), int(k), nid.data_ptr(), nt.data_ptr(), nx.data_ptr() if self.D else None,
            _cabi.current_stream(self.device)))
        return nid, nt, nx

    def sample_ids(self, seeds: Tensor, tq: Tensor, cut: Tensor, k: int, B: int,
                   cut_group: int = 1):
        """`sample` without the feature block (tgm_csr_sample_ids): (nbr_nids, nbr_edge_time, eid)."""
        S, dev = seeds.numel(), self.device
        nid = torch.empty((S, k), dtype=torch.int32, device=dev)
        nt = torch.empty((S, k), dtype=torch.int64, device=dev)
        eid = torch.empty((S, k), dtype=torch.int32, device=dev)
        _cabi.check(_cabi.lib.tgm_csr_sample_ids(
            self._handle, seeds.data_ptr(), tq.data_ptr(), cut.data_ptr(), int(cut_group), S,
            int(B), int(k), nid.data_ptr(), nt.data_ptr(), eid.data_ptr(),
            _cabi.current_stream(dev)))
        return nid, nt, eid

    def sample_edges(self, e_lo: int, e_hi: int, k: int, B: int, out=None):
        """Hop 0 for seeds = endpoints of stream edges [e_lo, e_hi) (src rows then dst rows per
        batch): no seed arrays, no search -- the prebuilt anchor table gives the window."""
        S = 2 * (e_hi - e_lo)
        nid, nt, nx = out if out is not None else self._alloc(S, k)
        _cabi.check(_cabi.lib.tgm_csr_sample_edges(
            self._handle, int(e_lo), int(e_hi), int(B), int(k), nid.data_ptr(), nt.data_ptr(),
            nx.data_ptr() if self.D else None, _cabi.current_stream(self.device)))
        return nid, nt, nx

    def sample_edges_host(self, e_lo: int, e_hi: int, k: int, B: int, host_in, host_out,
                          slot: int = 0, stream: Optional[int] = None) -> None:
        """Host-buffer form (tgm_csr_sample_edges_host): `host_in` = (src, dst, t, x) CPU tensors
        of the slab [e_lo, e_hi) (entries may be None), `host_out` = (nid, t, x) CPU tensors
        receiving the result; pinned tensors keep the copies asynchronous.  Stream-ordered: the
        outputs are valid after the stream is synchronised."""
        src, dst, t, x = host_in
        nid, nt, nx = host_out
        _cabi.check(_cabi.lib.tgm_csr_sample_edges_host(
            self._handle, int(e_lo), int(e_hi), int(B), int(k), _cabi.ptr(src), _cabi.ptr(dst),
            _cabi.ptr(t), _cabi.ptr(x), nid.data_ptr(), nt.data_ptr(),
            nx.data_ptr() if self.D else None, int(slot),
            _cabi.current_stream(self.device) if stream is None else stream))

    def sample_edges_ids(self, e_lo: int, e_hi: int, k: int, B: int, search: bool = False,
                         out=None):
        """Hop 0 without the feature block (tgm_csr_sample_edges_ids): (nbr_nids, nbr_edge_time,
        eid) with nbr_edge_x[s, c] == edge_x[eid[s, c]] (zeros where eid == -1)."""
        S = 2 * (e_hi - e_lo)
        if out is None:
            out = (torch.empty((S, k), dtype=torch.int32, device=self.device),
                   torch.empty((S, k), dtype=torch.int64, device=self.device),
                   torch.empty((S, k), dtype=torch.int32, device=self.device))
        _cabi.check(_cabi.lib.tgm_csr_sample_edges_ids(
            self._handle, int(e_lo), int(e_hi), int(B), int(k), int(search), _cabi.ptr(out[0]),
            _cabi.ptr(out[1]), _cabi.ptr(out[2]), _cabi.current_stream(self.device)))
        return out

    def sample_edges_mean(self, e_lo: int, e_hi: int, k: int, B: int, search: bool = False,
                          with_ids: bool = True, out=None):
        """Fused hop-0 sample + masked mean (tgm_csr_sample_edges_mean): (nbr_nids, nbr_edge_time,
        mean[S, D]); the first two are None when `with_ids` is False."""
        S = 2 * (e_hi - e_lo)
        if out is None:
            out = (torch.empty((S, k), dtype=torch.int32, device=self.device) if with_ids else None,
                   torch.empty((S, k), dtype=torch.int64, device=self.device) if with_ids else None,
                   torch.empty((S, self.D), dtype=torch.float32, device=self.device))
        _cabi.check(_cabi.lib.tgm_csr_sample_edges_mean(
            self._handle, int(e_lo), int(e_hi), int(B), int(k), int(search), _cabi.ptr(out[0]),
            _cabi.ptr(out[1]), None, out[2].data_ptr(), _cabi.current_stream(self.device)))
        return out

    def sample_edges_host_ids(self, e_lo: int, e_hi: int, k: int, B: int, host_in, host_out,
                              slot: int = 0, stream: Optional[int] = None) -> None:
        """tgm_csr_sample_edges_host_ids: `host_in` = (src, dst, t) CPU tensors of the slab,
        `host_out` = (nid, t, eid) CPU tensors; 16 bytes per sampled slot come back."""
        src, dst, t = host_in[:3]
        nid, nt, eid = host_out
        _cabi.check(_cabi.lib.tgm_csr_sample_edges_host_ids(
            self._handle, int(e_lo), int(e_hi), int(B), int(k), _cabi.ptr(src), _cabi.ptr(dst),
            _cabi.ptr(t), nid.data_ptr(), nt.data_ptr(), eid.data_ptr(), int(slot),
            _cabi.current_stream(self.device) if stream is None else stream))

    def sample_edges_host_mean(self, e_lo: int, e_hi: int, k: int, B: int, host_in, host_out,
                               slot: int = 0, stream: Optional[int] = None) -> None:
        """tgm_csr_sample_edges_host_mean: `host_out` = (nid | None, t | None, mean[S, D])."""
        src, dst, t = host_in[:3]
        nid, nt, mean = host_out
        _cabi.check(_cabi.lib.tgm_csr_sample_edges_host_mean(
            self._handle, int(e_lo), int(e_hi), int(B), int(k), _cabi.ptr(src), _cabi.ptr(dst),
            _cabi.ptr(t), _cabi.ptr(nid), _cabi.ptr(nt), mean.data_ptr(), int(slot),
            _cabi.current_stream(self.device) if stream is None else stream))

    # -- whole windows ----------------------------------------------------------------------
    def window_seed_tensors(self, e_lo: int, e_hi: int, neg: Optional[Tensor] = None):
        """hop-0 seeds/times/cuts of the window laid out batch by batch in the reference's
        seed order [src | dst (| neg)] (recency.py:181-233)."""
        st, bs = self._storage, self.batch_size
        src, dst, t = st._src[e_lo:e_hi], st._dst[e_lo:e_hi], st._t[e_lo:e_hi]
        n = e_hi - e_lo
        parts = [src, dst] + ([neg] if neg is not None else [])
        P = len(parts)
        nfull, rem = divmod(n, bs)

        def lay(cols: Sequence[Tensor]) -> Tensor:
            full = torch.stack([c[:nfull * bs].view(nfull, bs) for c in cols], 1).reshape(-1)
            if rem:
                return torch.cat([full] + [c[nfull * bs:] for c in cols])
            return full

        seeds = lay(parts)
        times = lay([t] * P)
        starts = e_lo + torch.arange(0, n, bs, device=self.device, dtype=torch.int64)
        counts = torch.full((len(starts),), bs * P, device=self.device, dtype=torch.int64)
        if rem:
            counts[-1] = rem * P
        cut = torch.repeat_interleave(starts, counts)
        return seeds.contiguous(), times.contiguous(), cut

    def sample_window(self, e_lo: int, e_hi: int, num_nbrs: Sequence[int],
                      neg: Optional[Tensor] = None, lazy_edge_x: bool = False) -> List[HopSample]:
        """All hops for the loader batches covering stream edges [e_lo, e_hi).

        Rows are the concatenation over batches of what RecencyNeighborHook puts on each batch;
        `split_window` cuts them back into per-batch views.  `lazy_edge_x`: the hops carry
        `LazyEdgeRows` (edge ids into the store's feature table) instead of feature blocks."""
        B = max(num_nbrs)
        hops: List[HopSample] = []
        seeds, times, cut = self.window_seed_tensors(e_lo, e_hi, neg)
        group = 1
        lazy = lazy_edge_x and self.D > 0 and B <= 32
        for h, k in enumerate(num_nbrs):
            if lazy:
                if h == 0 and neg is None and not self.directed:
                    nid, nt, eid = self.sample_edges_ids(e_lo, e_hi, k, B)
                else:
                    nid, nt, eid = self.sample_ids(seeds, times, cut, k, B, cut_group=group)
                nx = LazyEdgeRows(self._storage._x, eid)
            elif h == 0 and neg is None:
                nid, nt, nx = self.sample_edges(e_lo, e_hi, k, B)
            else:
                nid, nt, nx = self.sample(seeds, times, cut, k, B, cut_group=group)
            hops.append(HopSample(seeds, times, nid, nt, nx))
            seeds, times = nid.reshape(-1), nt.reshape(-1)
            group *= k
        return hops

    def split_window(self, hops: List[HopSample], e_lo: int, e_hi: int, seeds_per_edge: int = 2):
        """Yield (batch_lo, batch_hi, [HopSample views]) per loader batch of the window."""
        bs = self.batch_size
        row = 0
        for lo in range(e_lo, e_hi, bs):
            hi = min(lo + bs, e_hi)
            n = (hi - lo) * seeds_per_edge
            views, a, b = [], row, row + n
            for hop in hops:
                k = hop.nbr_nids.shape[1]
                views.append(HopSample(hop.seed_nids[a:b], hop.seed_times[a:b], hop.nbr_nids[a:b],
                                       hop.nbr_edge_time[a:b], hop.nbr_edge_x[a:b]))
                a, b = a * k, b * k
            yield lo, hi, views
            row += n


def reference_rng_picks(counts, k: int):
    """Candidate ordinals the reference would keep: `counts[i]` candidates for the i-th unique
    seed node (ascending node order, array_backend.py:118,:147); a node with more than k keeps
    `random.sample(range(count), k)` -- the same call on CPython's GLOBAL generator as upstream
    (:152-153: sampling a list selects by position, so the stream of draws is identical) -- the
    others keep all their candidates in order.  Returns a list of k-long lists, -1 = padding."""
    import random
    picks = []
    for c in counts:
        if c > k:
            picks.append(random.sample(range(c), k))
        else:
            picks.append(list(range(c)) + [-1] * (k - c))
    return picks


def full_history_neighbors(storage, seed_nodes: Tensor, num_nbrs: int, slice, directed: bool,
                           rng_seed: Optional[int] = None, reference_rng: bool = False):
    """DGStorageArrayBackend.get_nbrs (array_backend.py:108-171) on the device: neighbours among
    all edges of `slice`, left-aligned / right-padded.  The (edge, side)-ordered adjacency is
    built once per `directed` flag and cached on the storage."""
    cache = storage._node_cache
    key = ('uniform_csr', bool(directed))
    if key not in cache:
        cache[key] = RecencyCSR(storage, 1, directed=directed, colocate_x=False)
    csr = cache[key]
    if rng_seed is None and not reference_rng:  # a fresh draw per call, reproducible under torch.manual_seed
        rng_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    # a slice bounded by times only (what the hook passes: end_time = min(batch time) - 1) is handed
    # to the kernel as it is; only index-bounded slices are resolved to an edge range first
    by_time = (not reference_rng and getattr(slice, 'start_idx', None) is None
               and getattr(slice, 'end_idx', None) is None)
    lo, hi = (0, 0) if by_time else storage.edge_range(slice)
    dev = storage.device
    seeds = seed_nodes.to(device=dev, dtype=torch.int32).contiguous()
    S, D = seeds.numel(), csr.D
    nid = torch.empty((S, num_nbrs), dtype=torch.int32, device=dev)
    nt = torch.empty((S, num_nbrs), dtype=torch.int64, device=dev)
    nx = torch.empty((S, num_nbrs, D), dtype=torch.float32, device=dev)
    if reference_rng and S:
        # bit-exact with the reference under `random.seed`: counts to the host (one sync), the
        # reference's own random.sample calls there, the gather on the device
        uniq, inverse = torch.unique(seeds, return_inverse=True)  # ascending, like upstream (:118)
        counts = torch.empty((uniq.numel(),), dtype=torch.int64, device=dev)
        _cabi.check(_cabi.lib.tgm_csr_candidate_counts(
            csr.handle, uniq.data_ptr(), uniq.numel(), lo, hi, counts.data_ptr(),
            _cabi.current_stream(dev)))
        picks = torch.tensor(reference_rng_picks(counts.cpu().tolist(), int(num_nbrs)),
                             dtype=torch.int32).reshape(-1, num_nbrs).to(dev)[inverse].contiguous()
        _cabi.check(_cabi.lib.tgm_csr_gather_picks(
            csr.handle, seeds.data_ptr(), S, lo, hi, int(num_nbrs), picks.data_ptr(),
            nid.data_ptr(), nt.data_ptr(), nx.data_ptr() if D else None,
            _cabi.current_stream(dev)))
        return nid, nt, nx
    if by_time:
        t_lo, t_hi = slice.start_time, slice.end_time
        _cabi.check(_cabi.lib.tgm_csr_sample_uniform_time(
            csr.handle, seeds.data_ptr(), S, int(t_lo or 0), t_lo is not None, int(t_hi or 0),
            t_hi is not None, int(num_nbrs), int(rng_seed), nid.data_ptr(), nt.data_ptr(),
            nx.data_ptr() if D else None, _cabi.current_stream(dev)))
        return nid, nt, nx
    _cabi.check(_cabi.lib.tgm_csr_sample_uniform(
        csr.handle, seeds.data_ptr(), S, lo, hi, int(num_nbrs), int(rng_seed), nid.data_ptr(),
        nt.data_ptr(), nx.data_ptr() if D else None, _cabi.current_stream(dev)))
    return nid, nt, nx
