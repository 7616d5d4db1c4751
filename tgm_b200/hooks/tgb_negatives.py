"""Pre-generated TGB negatives as seed producers (tgm/hooks/negatives/tgb_sampler.py:16-309).

The candidate lists themselves come from the third-party `tgb` package (py-tgb: a dictionary
lookup per positive edge over a pickled evaluation set) and stay on the host; what runs per batch
on the device is what the reference does with them (tgb_sampler.py:92-134):
  neg            sorted unique ids over all candidate lists     (tgm_dedup_unique, one launch)
  neg_batch_list one int32 tensor per positive edge             (ONE host->device copy, views)
  neg_time       randint(t_min, t_max + 1) from a generator re-seeded with 0 on every call
so `neg` is a third of the neighbour sampler's seeds exactly as with the random sampler.

`tgb` is imported where the reference imports it (constructor); `neg_sampler=` injects any object
with `query_batch(...)` / `load_eval_set(...)` instead (an addition to the reference signature:
the package and its dataset files are absent from an offline box).
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, List, Optional

import numpy as np
import torch
from torch import Tensor

from tgm_b200.hooks.base import StatelessHook
from tgm_b200.hooks.hook_manager import register_hook_class


class TGBNegativeEdgeSamplerBase(StatelessHook):
    """Common part of the three TGB hooks: evaluation-set loading, the per-batch query and the
    batch attributes.  Subclasses name the sampler class and the query arguments."""

    _cls_requires = {'edge_src', 'edge_dst', 'edge_time'}
    _cls_produces = {'neg', 'neg_batch_list', 'neg_time'}
    _dataset_prefix: str = ''

    def __init__(self, dataset_name: str, split_mode: str, id: Optional[str] = None,
                 neg_sampler: Any = None) -> None:
        if split_mode not in ('val', 'test'):
            raise ValueError(f'split_mode must be "val" or "test", got: {split_mode}')
        if neg_sampler is None:
            try:
                from tgb.utils.info import DATA_VERSION_DICT, PROJ_DIR
            except ImportError:
                raise ImportError(
                    f'TGB required for {self.__class__.__name__}, try `pip install py-tgb`')
        if not dataset_name.startswith(f'{self._dataset_prefix}-'):
            raise ValueError(
                f'{self.__class__.__name__} should only be registered for '
                f'"{self._dataset_prefix}-xxx" datasets, but got: {dataset_name}')
        if neg_sampler is None:
            neg_sampler = self._build_sampler(dataset_name)
            version = DATA_VERSION_DICT.get(dataset_name, 1)
            suffix = f'_v{version}' if version > 1 else ''
            root = Path(PROJ_DIR + 'datasets') / dataset_name.replace('-', '_')
            neg_sampler.load_eval_set(
                fname=str(root / f'{dataset_name}_{split_mode}_ns{suffix}.pkl'),
                split_mode=split_mode)
        self.neg_sampler = neg_sampler
        self.split_mode = split_mode
        self._idset = None
        self._init_hook(id=id)

    def _build_sampler(self, dataset_name: str) -> Any:
        raise NotImplementedError

    def _query_batch(self, batch) -> list:
        raise NotImplementedError

    def _import_sampler(self, module: str, name: str):
        try:
            return getattr(__import__(module, fromlist=[name]), name)
        except ImportError:
            raise ImportError(
                f'TGB required for {self.__class__.__name__}, try `pip install py-tgb`')

    def _unique(self, flat: Tensor, dg) -> Tensor:
        """torch.unique(flat) for int32 ids: the bitmap kernel on the device, torch elsewhere."""
        if flat.is_cuda:
            from tgm_b200.hooks.dedup import _BatchIdSet
            n = int(dg.num_nodes)
            if self._idset is None or self._idset.num_nodes != n or \
                    self._idset.device != flat.device:
                self._idset = _BatchIdSet(n, flat.device)
            got = self._idset.unique([(flat, False)])
            if got is not None:
                return got.clone()  # the id set's output buffer is reused by the next batch
        return torch.unique(flat)

    def __call__(self, dg, batch):
        dev = dg.device
        E = batch.edge_src.size(0)
        if E == 0:
            neg = torch.empty((0,), dtype=torch.int32, device=dev)
            neg_time = torch.empty((0,), dtype=torch.int64, device=dev)
            per_edge: List[Tensor] = []
        else:
            try:
                lists = self._query_batch(batch)
            except ValueError as e:
                raise ValueError(
                    f'{self._dataset_prefix.upper()} Negative sampling failed for split_mode='
                    f'{self.split_mode}. Try updating your TGB package: '
                    '`pip install --upgrade py-tgb`') from e
            # every candidate list in one pinned-size host array, one copy, then views
            # (the reference builds E separate device tensors, tgb_sampler.py:109-112)
            arrays = [np.asarray(x, dtype=np.int32).reshape(-1) for x in lists]
            sizes = [a.size for a in arrays]
            flat = torch.from_numpy(np.concatenate(arrays) if arrays else
                                    np.empty(0, np.int32)).to(dev)
            per_edge = list(flat.split(sizes)) if sizes else []
            neg = self._unique(flat, dg)
            # a fresh generator seeded with 0 on every call: the same draws for every batch of
            # equal size and range (tgb_sampler.py:120-129)
            gen = torch.Generator(device=dev)
            gen.manual_seed(0)
            neg_time = torch.randint(int(batch.edge_time.min().item()),
                                     int(batch.edge_time.max().item()) + 1, (neg.size(0),),
                                     device=dev, generator=gen)
        self.add_batch_attribute(batch, 'neg', neg)
        self.add_batch_attribute(batch, 'neg_batch_list', per_edge)
        self.add_batch_attribute(batch, 'neg_time', neg_time)
        return batch


@register_hook_class
class TGBNegativeEdgeSamplerHook(TGBNegativeEdgeSamplerBase):
    """tgbl-* datasets: NegativeEdgeSampler.query_batch(src, dst, t, split_mode=...)."""

    # (the hook base folds the LEAF class's sets into the instance, as upstream: every leaf names both)
    _cls_requires = {'edge_src', 'edge_dst', 'edge_time'}
    _cls_produces = {'neg', 'neg_batch_list', 'neg_time'}
    _dataset_prefix = 'tgbl'

    def _build_sampler(self, dataset_name: str) -> Any:
        cls = self._import_sampler('tgb.linkproppred.negative_sampler', 'NegativeEdgeSampler')
        return cls(dataset_name=dataset_name)

    def _query_batch(self, batch) -> list:
        return self.neg_sampler.query_batch(batch.edge_src, batch.edge_dst, batch.edge_time,
                                            split_mode=self.split_mode)


@register_hook_class
class TGBTHGNegativeEdgeSamplerHook(TGBNegativeEdgeSamplerBase):
    """thgl-* (heterogeneous) datasets: the query carries the edge type."""

    _cls_requires = {'edge_src', 'edge_dst', 'edge_time', 'edge_type'}
    _cls_produces = {'neg', 'neg_batch_list', 'neg_time'}
    _dataset_prefix = 'thgl'

    def __init__(self, dataset_name: str, split_mode: str, first_node_id: int, last_node_id: int,
                 node_type: Tensor, id: Optional[str] = None, neg_sampler: Any = None) -> None:
        if first_node_id < 0 or last_node_id < 0:
            raise ValueError('First and last ID of node must be positive')
        if node_type is None:
            raise ValueError('Node type must not be None')
        if node_type.shape[0] < last_node_id:
            raise ValueError(f'last_node_id {last_node_id} must be within node_type')
        self._first_node_id, self._last_node_id = first_node_id, last_node_id
        self._node_type = node_type
        super().__init__(dataset_name, split_mode, id, neg_sampler)

    def _build_sampler(self, dataset_name: str) -> Any:
        cls = self._import_sampler('tgb.linkproppred.thg_negative_sampler',
                                   'THGNegativeEdgeSampler')
        return cls(dataset_name=dataset_name, first_node_id=self._first_node_id,
                   last_node_id=self._last_node_id, node_type=self._node_type.cpu().numpy())

    def _query_batch(self, batch) -> list:
        return self.neg_sampler.query_batch(batch.edge_src, batch.edge_dst, batch.edge_time,
                                            batch.edge_type, split_mode=self.split_mode)


@register_hook_class
class TGBTKGNegativeEdgeSamplerHook(TGBNegativeEdgeSamplerBase):
    """tkgl-* (knowledge graph) datasets: the query carries the edge type."""

    _cls_requires = {'edge_src', 'edge_dst', 'edge_time', 'edge_type'}
    _cls_produces = {'neg', 'neg_batch_list', 'neg_time'}
    _dataset_prefix = 'tkgl'

    def __init__(self, dataset_name: str, split_mode: str, first_dst_id: int, last_dst_id: int,
                 id: Optional[str] = None, neg_sampler: Any = None) -> None:
        if first_dst_id < 0 or last_dst_id < 0:
            raise ValueError('First and last ID of node must be positive')
        self._first_dst_id, self._last_dst_id = first_dst_id, last_dst_id
        super().__init__(dataset_name, split_mode, id, neg_sampler)

    def _build_sampler(self, dataset_name: str) -> Any:
        cls = self._import_sampler('tgb.linkproppred.tkg_negative_sampler',
                                   'TKGNegativeEdgeSampler')
        return cls(dataset_name=dataset_name, first_dst_id=self._first_dst_id,
                   last_dst_id=self._last_dst_id)

    def _query_batch(self, batch) -> list:
        return self.neg_sampler.query_batch(batch.edge_src, batch.edge_dst, batch.edge_time,
                                            batch.edge_type, split_mode=self.split_mode)
