"""NeighborSamplerHook: drop-in for tgm/hooks/neighbors/uniform.py:20-210 on the B200.

Stateless uniform sampling over the full history before the batch: per hop one
`storage.get_nbrs(seeds, k, DGSliceTracker(end_time=min(batch.edge_time) - 1), directed)` call
(uniform.py:122-127), served by `tgm_csr_sample_uniform` instead of a Python loop over every
edge of the history (array_backend.py:125-137).  Outputs are right-padded, per-seed query times
are ignored (only the batch-min cut applies) -- both as in the reference.

How the device serves a call.  The store keeps one extra adjacency for this hook, built lazily
and cached on the storage object: `RecencyCSR(storage, batch_size=1, colocate_x=False)`, i.e. every
node's incident entries {neighbour, edge, time} in (edge, side) order -- the order in which
get_nbrs appends candidates while it walks the history (array_backend.py:132-137).  A slice of
the history [e_lo, e_hi) is then, per seed node, a contiguous run of that node's entries found
by two binary searches over the edge index; nothing is scanned.  With at most k candidates the
run is copied out left-aligned (the reference's rows, bit for bit).  With more than k:

* default: `tgm_csr_sample_uniform` draws a uniform k-subset per node with Floyd's algorithm from
  a counter-based generator keyed by (call seed, node id), so all occurrences of a node in one
  call share the draw exactly as the reference's `inverse_indices == i` write does (:166);
* `reference_rng=True`: `tgm_csr_candidate_counts` returns the run lengths of the unique seed
  nodes, the host makes the reference's own `random.sample` calls in ascending node order
  (`tgm_b200.sampler.reference_rng_picks`), and `tgm_csr_gather_picks` gathers those candidate
  ordinals: bit-exact under `random.seed`, one host synchronisation per hop.

Seed validation (missing / None / non-tensor / wrong-rank attributes, negative ids or times)
raises the reference's exceptions; the bounds checks of all seed keys share one host read.
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from tgm_b200.core.storage import DGSliceTracker
from tgm_b200.hooks.base import SeedableHook, StatelessHook
from tgm_b200.hooks.hook_manager import register_hook_class


@register_hook_class
class NeighborSamplerHook(StatelessHook, SeedableHook):
    """Load neighbors of each seed node by uniform sampling over its history."""

    _cls_requires = {'edge_src', 'edge_dst', 'edge_time'}
    _cls_produces = {'seed_nids', 'seed_times', 'nbr_nids', 'nbr_edge_time', 'nbr_edge_x',
                     'seed_node_nbr_mask'}

    def __init__(self, num_nbrs: List[int], seed_nodes_keys: List[str],
                 seed_times_keys: List[str], directed: bool = False,
                 id: Optional[str] = None, reference_rng: bool = False) -> None:
        """`reference_rng=True` (an addition to the reference signature): sub-sample with the
        reference's own `random.sample` stream -- bit-exact outputs under `random.seed`, one host
        sync per hop -- instead of the device generator."""
        if not len(num_nbrs):
            raise ValueError('num_nbrs must be non-empty')
        if not all(isinstance(x, int) and x > 0 for x in num_nbrs):
            raise ValueError('Each value in num_nbrs must be a positive integer')
        if len(seed_nodes_keys) != len(seed_times_keys):
            raise ValueError(
                f'len(seed_nodes_keys) ({len(seed_nodes_keys)}) != len(seed_times_keys) '
                f'({len(seed_times_keys)})\nseed_nodes_keys={seed_nodes_keys}, '
                f'seed_times_keys={seed_times_keys}')
        self._num_nbrs = num_nbrs
        self._directed = directed
        self._reference_rng = bool(reference_rng)
        self._seed_nodes_keys = seed_nodes_keys
        self._seed_times_keys = seed_times_keys
        self._warned_seed_None = False
        self._init_hook(id=id, seed_keys=seed_nodes_keys)

    @property
    def num_nbrs(self) -> List[int]:
        return self._num_nbrs

    def __call__(self, dg, batch):
        seeds_out: List[Tensor] = []
        times_out: List[Tensor] = []
        nids: List[Tensor] = []
        nts: List[Tensor] = []
        nxs: List[Tensor] = []
        seed_nodes, seed_times, seed_mask = self._get_seed_tensors(batch)
        if not seed_nodes.numel():
            for _ in self._num_nbrs:  # uniform.py:93-101
                seeds_out.append(torch.empty(0, dtype=torch.int32))
                times_out.append(torch.empty(0, dtype=torch.int64))
                nids.append(torch.empty(0, dtype=torch.int32))
                nts.append(torch.empty(0, dtype=torch.int64))
                nxs.append(torch.empty(0, dg.edge_x_dim).float())
        else:
            cut = DGSliceTracker(end_time=self._min_edge_time(dg, batch) - 1)  # uniform.py:125
            for hop, k in enumerate(self._num_nbrs):
                if hop > 0:
                    seed_nodes = nids[hop - 1].flatten()
                    seed_times = nts[hop - 1].flatten()
                extra = {'reference_rng': True} if self._reference_rng else {}
                nid, nt, nx = dg._storage.get_nbrs(seed_nodes, num_nbrs=k, slice=cut,
                                                   directed=self._directed, **extra)
                seeds_out.append(seed_nodes)
                times_out.append(seed_times)
                nids.append(nid)
                nts.append(nt)
                nxs.append(nx)
        self.add_batch_attribute(batch, 'seed_nids', seeds_out)
        self.add_batch_attribute(batch, 'seed_times', times_out)
        self.add_batch_attribute(batch, 'nbr_nids', nids)
        self.add_batch_attribute(batch, 'nbr_edge_time', nts)
        self.add_batch_attribute(batch, 'nbr_edge_x', nxs)
        self.add_batch_attribute(batch, 'seed_node_nbr_mask', seed_mask)
        return batch

    @staticmethod
    def _min_edge_time(dg, batch) -> int:
        """`batch.edge_time.min()` without the device round trip when the batch still carries the
        loader's own views of a device store: the stream is time-sorted, so the minimum is the
        first row's time, read from the store's host copy of the edge times."""
        slab = getattr(batch, '_slab', None)
        store = getattr(dg, '_storage', None)
        if slab is not None and slab[0] is store and batch.edge_time is slab[5] and slab[2] > slab[1]:
            host_t = getattr(store, '_edge_time_host', None)
            if host_t is not None and host_t.numel() > slab[1]:
                return int(host_t[slab[1]])
        return int(batch.edge_time.min())

    def _get_seed_tensors(self, batch) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
        """uniform.py:144-210 (no upper bound on ids: `_num_nodes = inf`, :184-185); the bounds
        checks of all keys share one host sync."""
        device = batch.edge_src.device
        seeds: List[Tensor] = []
        times: List[Tensor] = []
        mask: Dict[str, Tensor] = {}
        checks: List[Tuple[str, Tensor, bool]] = []
        offset = 0
        for node_attr, time_attr in zip(self._seed_nodes_keys, self._seed_times_keys):
            missing = [a for a in (node_attr, time_attr) if not hasattr(batch, a)]
            if missing:
                raise ValueError(f'Missing seed attributes {missing} on batch')
            for name, tensor in ((node_attr, getattr(batch, node_attr)),
                                 (time_attr, getattr(batch, time_attr))):
                if tensor is None:
                    if not self._warned_seed_None:
                        warnings.warn(
                            f'Seed attribute {name} is None on this batch, skipping this batch. '
                            'Future occurrences will also be skipped but the warning will be '
                            'suppressed', UserWarning)
                        self._warned_seed_None = True
                    break
                if not isinstance(tensor, Tensor):
                    raise ValueError(f'{name} must be a Tensor, got {type(tensor)}')
                if tensor.ndim != 1:
                    raise ValueError(f'{name} must be 1-D, got shape {tensor.shape}')
                if tensor.numel():
                    checks.append((name, tensor, name == node_attr))
                if name == node_attr:
                    seeds.append(tensor.to(device))
                    mask[name] = torch.arange(offset, offset + tensor.shape[0], device=device)
                    offset += tensor.shape[0]
                else:
                    times.append(tensor.to(device))
        if checks:
            lows = torch.stack([t.min().to(torch.int64) for _, t, _ in checks]).cpu().tolist()
            for (name, t, is_node), lo in zip(checks, lows):
                if lo < 0 and is_node:
                    raise ValueError(f'Seed nodes in {name} must satisfy 0 <= x < inf, '
                                     f'got values in range [{lo}, {int(t.max())}]')
                if lo < 0:
                    raise ValueError(f'Seed times in {name} must be >= 0, got min value: {lo}')
        if seeds and times:
            return torch.cat(seeds), torch.cat(times), mask
        return (torch.empty(0, dtype=torch.int32, device=device),
                torch.empty(0, dtype=torch.int64, device=device), mask)
