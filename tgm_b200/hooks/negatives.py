"""Negative destination sampling for link prediction.

RandomNegativeEdgeSamplerHook mirrors tgm/hooks/negatives/sampler.py:14-65: `neg` =
randint(low, high, round(neg_ratio * E_b)) int32 on dg.device, `neg_time` = edge_time copy.
It is a seed PRODUCER for the neighbour sampler, not part of the bandwidth-bound path
(SURVEY.md section 8f, N3): torch's device RNG is used as is.
"""
from __future__ import annotations

from typing import Optional

import torch

from tgm_b200.hooks.base import StatelessHook
from tgm_b200.hooks.hook_manager import register_hook_class


@register_hook_class
class RandomNegativeEdgeSamplerHook(StatelessHook):
    """Random negative destinations for dynamic link prediction (negative sampler, uniform)."""

    _cls_requires = {'edge_src', 'edge_dst', 'edge_time'}
    _cls_produces = {'neg', 'neg_time'}

    def __init__(self, low: int, high: int, neg_ratio: float = 1.0,
                 id: Optional[str] = None) -> None:
        if not 0 < neg_ratio <= 1:
            raise ValueError(f'neg_ratio must be in (0, 1], got: {neg_ratio}')
        if not low < high:
            raise ValueError(f'low ({low}) must be strictly less than high ({high})')
        self.low, self.high, self.neg_ratio = low, high, neg_ratio
        self._init_hook(id=id)

    def __call__(self, dg, batch):
        n = round(self.neg_ratio * batch.edge_dst.size(0))
        if n == 0:
            neg = torch.empty((0,), dtype=torch.int32, device=dg.device)
            neg_time = torch.empty((0,), dtype=torch.int64, device=dg.device)
        else:
            neg = torch.randint(self.low, self.high, (n,), dtype=torch.int32, device=dg.device)
            neg_time = batch.edge_time.clone()
        self.add_batch_attribute(batch, 'neg', neg)
        self.add_batch_attribute(batch, 'neg_time', neg_time)
        return batch
