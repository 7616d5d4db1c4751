"""Negative destination sampling for link prediction.

RandomNegativeEdgeSamplerHook mirrors tgm/hooks/negatives/sampler.py:14-65: `neg` =
randint(low, high, round(neg_ratio * E_b)) int32 on dg.device, `neg_time` = edge_time copy.

It is the seed PRODUCER of a third of the neighbour sampler's seeds, so
while the loader walks a device store front to back the negatives of a whole window of batches
are drawn by ONE launch (`tgm_negatives_window`) and each call hands out views; the window is
published on the batch (`batch._seed_windows`) so that RecencyNeighborHook can sample
[src | dst | neg] for all those batches in one launch per hop.

RNG stream (documented, and pinned by tests/test_gpu_negatives.py): batch j of a window drawn when
torch's CUDA generator stood at (seed, offset) gets exactly the numbers the j-th of consecutive
`torch.randint(low, high, (n,), dtype=torch.int32, device='cuda')` calls would draw -- the
reference's own device='cuda' stream -- and the generator is advanced by 4 per batch.  If the
walk stops early the generator is rewound to where per-batch calls would have left it.  Other
consumers of the generator between two batches (e.g. dropout) see a different offset than they
would after per-batch draws: the draws are the same kind of stream, not the same numbers.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
from torch import Tensor

from tgm_b200 import _cabi
from tgm_b200.core.storage import block_views
from tgm_b200.hooks.base import StatelessHook
from tgm_b200.hooks.hook_manager import register_hook_class


class SeedWindow(NamedTuple):
    """Seeds a producer hook drew ahead for the stream edges [e_lo, e_hi) of `store`:
    nodes[j] / times[j] belong to stream edge e_lo + j; every id lies in [low, high).
    `node_views[i]` / `time_views[i]` are the tensors handed to batch i of the window
    (batches of `batch_size` edges, the last one may be short): a consumer recognises an untouched
    batch attribute by identity."""
    store: object
    e_lo: int
    e_hi: int
    nodes: Tensor
    times: Tensor
    low: int
    high: int
    batch_size: int = 0
    node_views: tuple = ()
    time_views: tuple = ()


@register_hook_class
class RandomNegativeEdgeSamplerHook(StatelessHook):
    """Random negative destinations for dynamic link prediction (negative sampler, uniform)."""

    _cls_requires = {'edge_src', 'edge_dst', 'edge_time'}
    _cls_produces = {'neg', 'neg_time'}

    def __init__(self, low: int, high: int, neg_ratio: float = 1.0,
                 id: Optional[str] = None, window_batches: Optional[int] = None) -> None:
        """`window_batches` (an addition to the reference signature): how many batches are drawn
        per launch; None = 1024, 0 = one torch.randint per batch as upstream."""
        if not 0 < neg_ratio <= 1:
            raise ValueError(f'neg_ratio must be in (0, 1], got: {neg_ratio}')
        if not low < high:
            raise ValueError(f'low ({low}) must be strictly less than high ({high})')
        self.low, self.high, self.neg_ratio = low, high, neg_ratio
        self._window_batches = 1024 if window_batches is None else int(window_batches)
        self._win = None
        self._init_hook(id=id)

    def reset_state(self) -> None:
        self._leave_window()

    # -- window mode ------------------------------------------------------------------------
    def _leave_window(self) -> None:
        """Rewind the generator to where per-batch draws would have left it."""
        w, self._win = self._win, None
        if w is not None and w['served'] < w['batches']:
            gen = torch.cuda.default_generators[w['device'].index]
            if gen.get_offset() == w['offset_after']:  # nobody else drew in between
                gen.set_offset(w['offset'] + 4 * w['served'])

    def _windowed(self, dg, batch):
        slab = getattr(batch, '_slab', None)  # set by the loader: these rows of this store
        if slab is None:
            return None
        store, lo, hi = slab[0], slab[1], slab[2]
        w = self._win
        if w is not None and store is w['store'] and lo == w['next'] and \
                (hi - lo == w['bs'] or hi == w['e_hi']) and hi <= w['e_hi'] and \
                batch.edge_src is slab[3] and batch.edge_time is slab[5]:
            j = w['served']  # the common case: the next batch of the window
        else:
            # ranges of 2^28 and more take ATen's 64-bit draw: left to torch.randint itself
            if not self._window_batches or self.neg_ratio != 1.0 or \
                    self.high - self.low >= 1 << 28 or store is not getattr(dg, '_storage', None) or \
                    batch.edge_src is not slab[3] or batch.edge_time is not slab[5] or \
                    getattr(store, 'device', None) is None:
                return None
            if w is not None:
                self._leave_window()
            n = hi - lo
            if n > 65536:
                return None
            dev = store.device
            e_hi = min(lo + self._window_batches * n, store.num_edges)
            gen = torch.cuda.default_generators[dev.index]
            seed, offset = gen.initial_seed(), gen.get_offset()
            if offset % 4:
                return None
            total = e_hi - lo
            nodes = torch.empty((total,), dtype=torch.int32, device=dev)
            _cabi.check(_cabi.lib.tgm_negatives_window(
                seed & 0xFFFFFFFFFFFFFFFF, offset, self.low, self.high, n, total, nodes.data_ptr(),
                _cabi.current_stream(dev)))
            batches = -(-total // n)
            gen.set_offset(offset + 4 * batches)
            times = store._t[lo:e_hi].clone()  # `neg_time` is a copy upstream (sampler.py:63)
            nv, tv = block_views(nodes, n), block_views(times, n)
            w = self._win = {
                'store': store, 'device': dev, 'bs': n, 'next': lo, 'e_lo': lo, 'e_hi': e_hi,
                'offset': offset, 'offset_after': offset + 4 * batches, 'batches': batches,
                'served': 0, 'nodes': nv, 'times': tv,
                'pub': SeedWindow(store, lo, e_hi, nodes, times, self.low, self.high, n, nv, tv),
                'key': f'neg_{self._id}' if self._id else 'neg'}
            j = 0
        w['next'] = hi
        w['served'] = j + 1
        pubs = getattr(batch, '_seed_windows', None)
        if pubs is None:
            batch._seed_windows = {w['key']: w['pub']}
        else:
            pubs[w['key']] = w['pub']
        if j + 1 == w['batches']:
            self._win = None
        return w['nodes'][j], w['times'][j]

    def __call__(self, dg, batch):
        got = self._windowed(dg, batch)  # None for empty batches too (the loader marks no slab rows)
        if got is not None:
            neg, neg_time = got
        else:
            n = round(self.neg_ratio * batch.edge_dst.size(0))
            if n == 0:
                neg = torch.empty((0,), dtype=torch.int32, device=dg.device)
                neg_time = torch.empty((0,), dtype=torch.int64, device=dg.device)
            else:
                neg = torch.randint(self.low, self.high, (n,), dtype=torch.int32, device=dg.device)
                neg_time = batch.edge_time.clone()
        if self._id is None:
            batch.neg, batch.neg_time = neg, neg_time
        else:
            self.add_batch_attribute(batch, 'neg', neg)
            self.add_batch_attribute(batch, 'neg_time', neg_time)
        return batch
