"""DeduplicationHook: unique node ids of a batch + global->local mapping.

Mirrors tgm/hooks/dedup.py:12-67 (requires edge_src/edge_dst + seed keys; `nbr_nids*` keys
contribute their non-padded entries per hop; produces `unique_nids` (sorted) and
`global_to_local`).  The non-padded neighbour ids are gathered through the frontier-compaction
kernel (`tgm_frontier_compact`) instead of a boolean-mask index.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor

from tgm_b200 import _cabi
from tgm_b200.hooks.base import SeedableHook, StatelessHook
from tgm_b200.hooks.hook_manager import register_hook_class


def compact_frontier(nid: Tensor) -> Tensor:
    """Indices (int64, ascending) of the entries of flat int32 `nid` that are not padded."""
    nid = nid.contiguous()
    n = nid.numel()
    idx = torch.empty(n, dtype=torch.int64, device=nid.device)
    count = torch.empty(1, dtype=torch.int64, device=nid.device)
    _cabi.check(_cabi.lib.tgm_frontier_compact(nid.data_ptr(), n, idx.data_ptr(),
                                               count.data_ptr(),
                                               _cabi.current_stream(nid.device)))
    return idx[:int(count.item())]


@register_hook_class
class DeduplicationHook(StatelessHook, SeedableHook):
    """Unique nodes of a batch and a node ID mapper from global to batch-local ids."""

    _cls_requires = {'edge_src', 'edge_dst'}
    _cls_produces = {'unique_nids', 'global_to_local'}

    def __init__(self, seed_nodes_keys: Optional[List[str]] = None,
                 id: Optional[str] = None) -> None:
        self._init_hook(id=id, seed_keys=seed_nodes_keys)

    def __call__(self, dg, batch):
        device = batch.edge_src.device
        parts = [batch.edge_src, batch.edge_dst]
        for attr in self.requires:
            if not hasattr(batch, attr):
                raise ValueError(f'Missing seed node attribute {attr}')
            if 'nbr_nids' in attr:
                for hop_ids in getattr(batch, attr):
                    flat = hop_ids.flatten()
                    if flat.numel() and flat.is_cuda:
                        parts.append(flat[compact_frontier(flat)].to(device))
            else:
                value = getattr(batch, attr)
                if value is not None:
                    parts.append(value)
        unique = torch.unique(torch.cat(parts, 0), sorted=True)
        self.add_batch_attribute(batch, 'unique_nids', unique)
        self.add_batch_attribute(batch, 'global_to_local',
                                 lambda x: torch.searchsorted(unique, x).int())
        return batch
