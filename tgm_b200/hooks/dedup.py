"""DeduplicationHook: unique node ids of a batch + global->local mapping.

Mirrors tgm/hooks/dedup.py:12-67 (requires edge_src/edge_dst + seed keys; `nbr_nids*` keys
contribute their non-padded entries per hop; produces `unique_nids` (sorted) and
`global_to_local`).  The reference's mask-gather + cat + torch.unique (a sort) + searchsorted per
lookup are one bitmap pass on the device (`tgm_dedup_unique`: mark, popcount scan, emit; padded
slots dropped in flight) and an O(1) rank lookup (`tgm_dedup_map`); one host read per batch (the
unique count, which sizes the result -- torch.unique has the same one).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple

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


class _BatchIdSet:
    """Device state of one batch's id set (bitmap + popcount prefix); owned by the batch's
    `global_to_local` closure so an older batch's map stays valid."""

    _sizes_cache = {}

    def __init__(self, num_nodes: int, device: torch.device) -> None:
        self.num_nodes, self.device = num_nodes, device
        if num_nodes not in self._sizes_cache:
            w, p, b = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
            _cabi.check(_cabi.lib.tgm_dedup_sizes(num_nodes, ctypes.byref(w), ctypes.byref(p),
                                                  ctypes.byref(b)))
            self._sizes_cache[num_nodes] = (w.value, p.value, b.value)
        w, p, b = self._sizes_cache[num_nodes]
        # one allocation: [tmp | bitmap | prefix], every piece 16-byte aligned
        tmp_words = (b + 15) // 16 * 4
        self._buf = torch.empty(tmp_words + (w + 3) // 4 * 4 + p, dtype=torch.int32, device=device)
        self.tmp, self.tmp_bytes = self._buf[:tmp_words], b
        self.bitmap = self._buf[tmp_words:tmp_words + w]
        self.prefix = self._buf[tmp_words + (w + 3) // 4 * 4:]

    def unique(self, parts: List[Tuple[Tensor, bool]]) -> Optional[Tensor]:
        """Sorted unique ids of the int32 device arrays in `parts` [(ids, skip_padded)];
        None when an id lies outside [-1, num_nodes)."""
        n = len(parts)
        ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() if t.numel() else None for t, _ in parts])
        sizes = (ctypes.c_int64 * n)(*[t.numel() for t, _ in parts])
        skip = (ctypes.c_int32 * n)(*[int(s) for _, s in parts])
        total = sum(t.numel() for t, _ in parts)
        out = torch.empty(min(total, self.num_nodes) + 1, dtype=torch.int32, device=self.device)
        count = torch.empty(1, dtype=torch.int64, device=self.device)
        _cabi.check(_cabi.lib.tgm_dedup_unique(
            ptrs, sizes, skip, n, self.num_nodes, self.bitmap.data_ptr(), self.prefix.data_ptr(),
            self.tmp.data_ptr(), self.tmp_bytes, out.data_ptr(), count.data_ptr(),
            _cabi.current_stream(self.device)))
        c = int(count.item())
        return None if c < 0 else out[:c]

    def local(self, x: Tensor) -> Tensor:
        ids = x.to(device=self.device, dtype=torch.int32).contiguous()
        out = torch.empty(ids.shape, dtype=torch.int32, device=self.device)
        _cabi.check(_cabi.lib.tgm_dedup_map(
            self.bitmap.data_ptr(), self.prefix.data_ptr(), self.tmp.data_ptr(), self.num_nodes,
            ids.data_ptr(), ids.numel(), out.data_ptr(), _cabi.current_stream(self.device)))
        return out


@register_hook_class
class DeduplicationHook(StatelessHook, SeedableHook):
    """Unique nodes of a batch and a node ID mapper from global to batch-local ids."""

    _cls_requires = {'edge_src', 'edge_dst'}
    _cls_produces = {'unique_nids', 'global_to_local'}
    _MAX_PARTS = 8  # tgm_dedup_unique takes up to 8 id arrays per call

    def __init__(self, seed_nodes_keys: Optional[List[str]] = None,
                 id: Optional[str] = None) -> None:
        self._init_hook(id=id, seed_keys=seed_nodes_keys)

    def __call__(self, dg, batch):
        device = batch.edge_src.device
        if device.type != 'cuda':
            raise _cabi.TGMNativeError(-3, 'DeduplicationHook needs a CUDA graph (tgm_b200 has no '
                                           'CPU fallback)')
        parts = [(batch.edge_src, False), (batch.edge_dst, False)]
        for attr in self.requires:
            if not hasattr(batch, attr):
                raise ValueError(f'Missing seed node attribute {attr}')
            if 'nbr_nids' in attr:
                for hop_ids in getattr(batch, attr):
                    parts.append((hop_ids.flatten(), True))  # padded slots are dropped in flight
            else:
                value = getattr(batch, attr)
                if value is not None:
                    parts.append((value.flatten(), False))
        out_dtype = parts[0][0].dtype
        for t, _ in parts[1:]:
            out_dtype = torch.promote_types(out_dtype, t.dtype)  # what torch.cat would give
        parts = [(t.to(device=device, dtype=torch.int32).contiguous(), s) for t, s in parts
                 if t.numel()]
        if len(parts) > self._MAX_PARTS:
            for flag in (True, False):  # fold the surplus arrays of each kind into one
                same = [t for t, s in parts if s is flag]
                if len(same) > 1:
                    parts = [(t, s) for t, s in parts if s is not flag] + [(torch.cat(same), flag)]
        if not parts:
            parts = [(batch.edge_src.to(torch.int32), False)]
        store = getattr(dg, '_storage', None)
        num_nodes = max(1, int(getattr(store, 'num_nodes_global', 0) or 0))
        ids = _BatchIdSet(num_nodes, device)
        unique = ids.unique(parts)
        if unique is None:  # an id beyond the store's node range (e.g. user-supplied seeds)
            top = max(int(t.max().item()) for t, _ in parts)
            low = min(int(t.min().item()) for t, _ in parts)
            if low < -1:
                raise ValueError(f'negative node id {low} in the batch')
            ids = _BatchIdSet(top + 1, device)
            unique = ids.unique(parts)
        if unique.dtype != out_dtype:
            unique = unique.to(out_dtype)
        self.add_batch_attribute(batch, 'unique_nids', unique)
        self.add_batch_attribute(batch, 'global_to_local', ids.local)
        return batch
