"""CPU ORACLE (test infrastructure, NOT product code) for the recency neighbor-sampling hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  Nothing under `tgm_b200/` imports it: the product path is
the CUDA library and fails loudly without it.

Two independent restatements of the reference algorithm
(reference = tgm-team/tgm @ 5183dc9, `tgm/hooks/neighbors/recency.py`):

* `RingOracle`       -- stateful per-node circular buffers, numpy, one call per loader batch.
                        Follows the reference's state machine: `_get_recency_neighbors`
                        (recency.py:239-321) and `_update` (recency.py:323-399).
* `stateless_sample` -- pure-Python restatement that never keeps ring state: it derives every
                        answer from the immutable event stream and the batch boundaries
                        (SURVEY.md Appendix A.2).  Small cases only.

Parity pinning: both are checked against (a) the known-answer values asserted by the
reference's own unit tests (test/unit/test_hooks/test_recency_nbr_hook.py:344-885) restated in
tests/test_oracle_golden.py, and (b) fixtures under tests/golden/*.npz produced by running the
unmodified reference in the build container (tests/golden/make_golden.py).

Known reference defect reproduced on request: the int32 overflow of the update sort key
(recency.py:347-348).  `RingOracle(int32_key_overflow=True)` wraps the composite key the way
torch's int32*0-dim-int64 promotion does; the default (False) is the ideal semantics, equal to
the reference whenever num_nodes * (t_max + 1) < 2**31.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

PADDED_NODE_ID = -1  # tgm/constants.py:3


class RingOracle:
    """Stateful restatement of RecencyNeighborHook's ring buffers (recency.py:93-97)."""

    def __init__(self, num_nodes: int, num_nbrs: Sequence[int], edge_x_dim: int = 0,
                 directed: bool = False, int32_key_overflow: bool = False) -> None:
        if not len(num_nbrs):
            raise ValueError('num_nbrs must be non-empty')  # recency.py:66-67
        if not all(isinstance(x, int) and x > 0 for x in num_nbrs):
            raise ValueError('Each value in num_nbrs must be a positive integer')  # :68-69
        self.N = int(num_nodes)
        self.num_nbrs = list(num_nbrs)
        self.B = max(num_nbrs)  # recency.py:73
        self.D = int(edge_x_dim)
        self.directed = directed
        self.int32_key_overflow = int32_key_overflow
        self.reset_state()

    # recency.py:111-117
    def reset_state(self) -> None:
        self.ids = np.full((self.N, self.B), PADDED_NODE_ID, dtype=np.int32)
        self.times = np.zeros((self.N, self.B), dtype=np.int64)
        self.feats = np.zeros((self.N, self.B, self.D), dtype=np.float32)
        self.write_pos = np.zeros(self.N, dtype=np.int32)

    # recency.py:239-321
    def query(self, seeds: np.ndarray, tq: np.ndarray, k: int
              ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        seeds = np.asarray(seeds)
        tq = np.asarray(tq, dtype=np.int64)
        S, B = len(seeds), self.B
        rows = seeds.astype(np.int64)  # negative ids index from the end, as torch does (:256)
        ring_ids, ring_t = self.ids[rows], self.times[rows]
        wp = self.write_pos[rows].astype(np.int64)
        # unrolled order: column j is ring slot (wp - (B - j)) mod B, oldest .. newest (:263-264)
        slot = (wp[:, None] - np.arange(B, 0, -1)[None, :]) % B
        t_unrolled = np.take_along_axis(ring_t, slot, 1)
        id_unrolled = np.take_along_axis(ring_ids, slot, 1)
        ok = (t_unrolled < tq[:, None]) & (id_unrolled != PADDED_NODE_ID)  # :267-269
        last = np.where(ok.any(1), (ok * np.arange(B)).max(1), -1)  # :274-281
        # k-window ending at `last`; entries before position 0 are padding (:287-301)
        col = last[:, None] - np.arange(k - 1, -1, -1)[None, :]
        valid = col >= 0
        src_slot = np.take_along_axis(slot, np.clip(col, 0, None), 1)
        out_id = np.where(valid, np.take_along_axis(ring_ids, src_slot, 1), PADDED_NODE_ID)
        out_t = np.where(valid, np.take_along_axis(ring_t, src_slot, 1), 0)
        out_x = np.where(valid[:, :, None], self.feats[rows[:, None], src_slot],
                         np.float32(0.0)).astype(np.float32)  # :304-314
        return out_id.astype(np.int32), out_t.astype(np.int64), out_x.reshape(S, k, self.D)

    # recency.py:323-399
    def update(self, src: np.ndarray, dst: np.ndarray, t: np.ndarray,
               x: Optional[np.ndarray]) -> None:
        n = len(src)
        if x is None:
            x = np.zeros((n, self.D), dtype=np.float32)  # :325-329
        if self.directed:
            node, nbr, tt, xx = src, dst, t, x  # :331-336
        else:
            node = np.concatenate([src, dst])  # :339-342
            nbr = np.concatenate([dst, src])
            tt = np.concatenate([t, t])
            xx = np.concatenate([x, x])
        tt = tt.astype(np.int64)
        max_time = tt.max() + 1  # :347
        if self.int32_key_overflow:
            key = (node.astype(np.int64) * max_time).astype(np.int32).astype(np.int64) + tt
        else:
            key = node.astype(np.int64) * max_time + tt  # :348 (ideal semantics)
        perm = np.argsort(key, kind='stable')  # :349
        node, nbr, tt, xx = node[perm], nbr[perm], tt[perm], xx[perm]
        # runs of equal consecutive node ids (unique_consecutive, :364)
        starts = np.flatnonzero(np.r_[True, node[1:] != node[:-1]])
        counts = np.diff(np.r_[starts, len(node)])
        run = np.repeat(np.arange(len(starts)), counts)
        pos = np.arange(len(node)) - starts[run]
        keep = pos >= counts[run] - self.B  # keep last B per run (:373)
        node, nbr, tt, xx = node[keep], nbr[keep], tt[keep], xx[keep]
        starts = np.flatnonzero(np.r_[True, node[1:] != node[:-1]]) if len(node) else np.array([], int)
        counts = np.diff(np.r_[starts, len(node)])
        run = np.repeat(np.arange(len(starts)), counts)
        off = np.arange(len(node)) - starts[run]
        widx = (self.write_pos[node].astype(np.int64) + off) % self.B  # :389
        self.ids[node, widx] = nbr  # :392-394
        self.times[node, widx] = tt
        self.feats[node, widx, :] = xx
        np.add.at(self.write_pos, node.astype(np.int64), np.int32(1))  # :397-399

    def hook_call(self, seeds: np.ndarray, tq: np.ndarray, src: np.ndarray, dst: np.ndarray,
                  t: np.ndarray, x: Optional[np.ndarray]):
        """One `RecencyNeighborHook.__call__` (recency.py:119-171): all hops, then the push."""
        out = []
        if len(seeds):
            s, q = np.asarray(seeds, np.int32), np.asarray(tq, np.int64)
            for hop, k in enumerate(self.num_nbrs):
                if hop > 0:
                    s, q = out[-1][2].reshape(-1), out[-1][3].reshape(-1)  # :141-143
                nid, nt, nx = self.query(s, q, k)
                out.append((s, q, nid, nt, nx))
            if len(src):
                self.update(src, dst, t, x)  # :161-163
        return out


def stateless_sample(src: np.ndarray, dst: np.ndarray, t: np.ndarray, x: Optional[np.ndarray],
                     batch_size: int, num_nbrs: Sequence[int], seeds_per_batch,
                     directed: bool = False, start: int = 0):
    """Stateless restatement.  Pure Python, small inputs only.

    hist[v] is the list of entries (nbr, t, eid) of node v ordered by (batch, t, side, eid),
    side 0 = "v is the edge source".  A query (v, tq) issued by batch b sees the last
    B = max(num_nbrs) entries of hist[v] restricted to batches < b, keeps the ones up to
    the right-most entry with t < tq, and returns the last k of them right-aligned.

    seeds_per_batch(b, lo, hi) -> (seed ids, seed times) of batch b covering edges [lo, hi).
    Returns a list (one per batch) of per-hop tuples (seeds, times, nid, nt, nx).
    """
    E = len(src)
    D = 0 if x is None else x.shape[1]
    B = max(num_nbrs)
    hist: dict = {}
    results = []
    for b, lo in enumerate(range(start, E, batch_size)):
        hi = min(lo + batch_size, E)
        s, q = seeds_per_batch(b, lo, hi)
        s, q = np.asarray(s, np.int32), np.asarray(q, np.int64)
        hops = []
        for hop, k in enumerate(num_nbrs):
            if hop > 0:
                s, q = hops[-1][2].reshape(-1), hops[-1][3].reshape(-1)
            nid = np.full((len(s), k), PADDED_NODE_ID, np.int32)
            nt = np.zeros((len(s), k), np.int64)
            nx = np.zeros((len(s), k, D), np.float32)
            for i, (v, tq) in enumerate(zip(s.tolist(), q.tolist())):
                if v < 0:
                    continue  # row N-1 with tq == 0: nothing is < 0
                W = hist.get(v, [])[-B:]
                last = -1
                for j, e in enumerate(W):
                    if e[1] < tq:
                        last = j
                V = W[max(0, last + 1 - k):last + 1]
                for c, e in enumerate(V):
                    col = k - len(V) + c
                    nid[i, col], nt[i, col] = e[0], e[1]
                    if D:
                        nx[i, col] = x[e[2]]
            hops.append((s, q, nid, nt, nx))
        results.append(hops)
        ents = []
        for e in range(lo, hi):
            ents.append((int(src[e]), int(t[e]), 0, e, int(dst[e])))
            if not directed:
                ents.append((int(dst[e]), int(t[e]), 1, e, int(src[e])))
        ents.sort(key=lambda r: (r[0], r[1], r[2], r[3]))
        for v, tt, _, e, nb in ents:
            hist.setdefault(v, []).append((nb, tt, e))
    return results


def masked_mean(z: np.ndarray, nbr_nids: np.ndarray) -> np.ndarray:
    """examples/linkproppred/graphmixer.py:131-135: sum_k(z*mask) / clamp(sum mask, 1), fp32,
    neighbours accumulated left to right."""
    mask = (nbr_nids != PADDED_NODE_ID)
    acc = np.zeros((z.shape[0], z.shape[2]), np.float32)
    for c in range(z.shape[1]):
        acc = acc + z[:, c, :] * mask[:, c, None].astype(np.float32)
    cnt = np.clip(mask.sum(1, keepdims=True), 1, None).astype(np.float32)
    return (acc / cnt).astype(np.float32)


def time2vec(dt: np.ndarray, w: np.ndarray, b: np.ndarray, fused: bool = True) -> np.ndarray:
    """tgm/nn/modules/time_encoding.py:22-24: cos(Linear(1,d)(float32(dt))).  Returned in float64:
    the exact cosine of the float32 argument.  `fused` selects how Linear rounds x*w+b: once
    (FMA; torch's batched CPU GEMM in the build container) or twice (product, then sum; torch's
    single-row GEMV path).  The two coincide when b == 0 (the shipped init, :20)."""
    x = np.asarray(dt).astype(np.float32)[:, None]
    w32, b32 = np.asarray(w, np.float32)[None, :], np.asarray(b, np.float32)[None, :]
    if fused:  # the float64 product of two float32 is exact; one rounding to float32 at the end
        arg = (x.astype(np.float64) * w32.astype(np.float64) + b32.astype(np.float64)).astype(np.float32)
    else:
        arg = (x * w32).astype(np.float32) + b32
    return np.cos(arg.astype(np.float64))


def uniform_candidates(src: np.ndarray, dst: np.ndarray, e_lo: int, e_hi: int, seeds,
                       directed: bool = False):
    """Candidate lists of DGStorageArrayBackend.get_nbrs (array_backend.py:125-137): for every
    seed the (edge, neighbour) pairs of the edges [e_lo, e_hi) it touches, in edge order, the
    src-side entry before the dst-side entry of the same edge."""
    want = set(int(v) for v in np.asarray(seeds).reshape(-1))
    cand = {v: [] for v in want}
    for e in range(e_lo, e_hi):
        s, d = int(src[e]), int(dst[e])
        if s in cand:
            cand[s].append((e, d))
        if not directed and d in cand:
            cand[d].append((e, s))
    return cand


def uniform_sample_deterministic(src, dst, t, x, e_lo: int, e_hi: int, seeds, k: int,
                                 directed: bool = False):
    """get_nbrs (array_backend.py:108-171) for the seeds whose candidate count is <= k (no random
    sub-sampling involved): left-aligned, right-padded (-1, 0, 0.0).  Returns (nid, nt, nx,
    exact) where exact[i] is False for seeds with more than k candidates (rows left as padding;
    the caller checks those for set-validity instead)."""
    seeds = np.asarray(seeds).reshape(-1)
    D = 0 if x is None else x.shape[1]
    cand = uniform_candidates(src, dst, e_lo, e_hi, seeds, directed)
    S = len(seeds)
    nid = np.full((S, k), PADDED_NODE_ID, np.int32)
    nt = np.zeros((S, k), np.int64)
    nx = np.zeros((S, k, D), np.float32)
    exact = np.ones(S, bool)
    for i, v in enumerate(seeds.tolist()):
        c = cand[int(v)]
        if len(c) > k:
            exact[i] = False
            continue
        for j, (e, nb) in enumerate(c):
            nid[i, j], nt[i, j] = nb, t[e]
            if D:
                nx[i, j] = x[e]
    return nid, nt, nx, exact


def uniform_sample_reference_rng(src, dst, t, x, e_lo: int, e_hi: int, seeds, k: int,
                                 directed: bool = False):
    """get_nbrs (array_backend.py:108-171) INCLUDING the sub-sampling: the unique seed nodes are
    visited in ascending order (:118, :147) and a node with more than k candidates keeps
    `random.sample(candidates, k)` -- CPython's GLOBAL generator, so the caller controls the stream
    with `random.seed` exactly as a user of the reference does; all occurrences of a node share
    the draw (:166).  Left-aligned, right-padded (-1, 0, 0.0)."""
    import random
    seeds = np.asarray(seeds).reshape(-1)
    D = 0 if x is None else x.shape[1]
    cand = uniform_candidates(src, dst, e_lo, e_hi, seeds, directed)
    S = len(seeds)
    nid = np.full((S, k), PADDED_NODE_ID, np.int32)
    nt = np.zeros((S, k), np.int64)
    nx = np.zeros((S, k, D), np.float32)
    for v in sorted(set(int(u) for u in seeds.tolist())):
        c = cand[v]
        if not c:
            continue
        if len(c) > k:
            c = random.sample(c, k)
        rows = np.flatnonzero(seeds == v)
        for j, (e, nb) in enumerate(c):
            nid[rows, j], nt[rows, j] = nb, t[e]
            if D:
                nx[rows, j] = x[e]
    return nid, nt, nx


def reference_rng_picks(counts, k: int) -> np.ndarray:
    """The same draws as candidate ORDINALS: counts[i] = number of candidates of the i-th unique
    seed node (ascending node order); returns int32 [len(counts), k], -1 = padding.
    `random.sample` selects by position only, so sampling range(c) consumes the generator exactly
    like sampling the candidate list."""
    import random
    picks = np.full((len(counts), k), -1, np.int32)
    for i, c in enumerate(counts):
        c = int(c)
        if c > k:
            picks[i] = random.sample(range(c), k)
        elif c:
            picks[i, :c] = np.arange(c)
    return picks
