"""Fused gather -> attention: the sampler hands out edge ids instead of (S, k, D)
feature blocks (tgm_csr_sample_ids / RecencyNeighborHook(lazy_edge_x=True) -> LazyEdgeRows) and
TemporalAttention reads the rows in place from the store's table (tgm_attn_forward_rows).  Same
numbers as the materialised path: ids/times bit-exact against the reference fixtures, attention
output bit-identical to the dense-block kernel."""
import numpy as np
import pytest
import torch

from tests._golden import Golden, golden_files, golden_ids

pytestmark = pytest.mark.gpu

from tgm_b200 import (DGData, DGDataLoader, DGraph, HookManager, RecencyCSR,  # noqa: E402
                      RecencyNeighborHook)
from tgm_b200.nn import TGAT  # noqa: E402
from tgm_b200.sampler import LazyEdgeRows  # noqa: E402

DEV = 'cuda:0'


def _cases():
    out, names = [], []
    for p, n in zip(golden_files(), golden_ids()):
        g = Golden(p)
        if g.x is not None and max(g.num_nbrs) <= 32 and g.neg is None:
            out.append(p)
            names.append(n)
    return out, names


@pytest.mark.parametrize('path', _cases()[0], ids=_cases()[1])
def test_lazy_hook_matches_reference_fixture(path):
    g = Golden(path)
    ei = torch.from_numpy(np.stack([g.src, g.dst], 1).astype(np.int32))
    dg = DGraph(DGData.from_raw(torch.from_numpy(g.t), ei, torch.from_numpy(g.x)), device=DEV)
    hook = RecencyNeighborHook(num_nodes=g.N, num_nbrs=g.num_nbrs,
                               seed_nodes_keys=['edge_src', 'edge_dst'],
                               seed_times_keys=['edge_time', 'edge_time'], directed=g.directed,
                               window_batches=5, lazy_edge_x=True)
    hm = HookManager(keys=['g'])
    hm.register('g', hook)
    with hm.activate('g'):
        for ep in range(g.epochs):
            for b, batch in enumerate(DGDataLoader(dg, batch_size=g.bs, hook_manager=hm)):
                for h in range(len(g.num_nbrs)):
                    want = g.expect(ep, b, h)
                    nx = batch.nbr_edge_x[h]
                    assert isinstance(nx, LazyEdgeRows) and nx.shape == want[4].shape
                    assert np.array_equal(batch.nbr_nids[h].cpu().numpy(), want[2])
                    assert np.array_equal(batch.nbr_edge_time[h].cpu().numpy(), want[3])
                    assert np.array_equal(nx.materialize().cpu().numpy(), want[4])
                    assert np.array_equal((nx.rows < 0).cpu().numpy(), want[2] < 0)
            if ep + 1 < g.epochs:
                hm.reset_state()


def _graph(seed, N, E, T, D):
    rng = np.random.default_rng(seed)
    src, dst = rng.integers(0, N, E).astype(np.int32), rng.integers(0, N, E).astype(np.int32)
    t = np.sort(rng.integers(0, T, E)).astype(np.int64)
    x = rng.standard_normal((E, D)).astype(np.float32)
    return DGraph(DGData.from_raw(torch.from_numpy(t), torch.from_numpy(np.stack([src, dst], 1)),
                                  torch.from_numpy(x)), device=DEV)


def test_general_seed_ids_kernel_equals_the_row_kernel():
    dg = _graph(1, 400, 30000, 2000, 12)
    csr = RecencyCSR(dg._storage, 100)
    g = torch.Generator(device=DEV).manual_seed(0)
    S = 5000
    seeds = torch.randint(-1, 400, (S,), generator=g, device=DEV, dtype=torch.int32)
    tq = torch.randint(0, 2000, (S,), generator=g, device=DEV)
    cut = torch.randint(0, 300, (S,), generator=g, device=DEV) * 100
    nid, nt, nx = csr.sample(seeds, tq, cut, 7, 9)
    nid2, nt2, eid = csr.sample_ids(seeds, tq, cut, 7, 9)
    assert torch.equal(nid, nid2) and torch.equal(nt, nt2)
    assert torch.equal(LazyEdgeRows(dg._storage._x, eid).materialize(), nx)
    assert torch.equal(eid < 0, nid < 0)


@pytest.mark.parametrize('D,node_dim', [(172, 1), (16, 8), (6, 3)], ids=['wiki', 'vec', 'scalar_rows'])
def test_tgat_on_lazy_rows_is_bit_identical_to_the_dense_blocks(D, node_dim):
    N, E, bs, k = 900, 12000, 200, 10
    dg = _graph(2, N, E, 100000, D)
    torch.manual_seed(0)
    model = TGAT(node_dim=node_dim, edge_dim=D, time_dim=20, embed_dim=32, num_layers=2,
                 n_heads=2, dropout=0.0).to(DEV).eval()
    node_x = torch.randn(N, node_dim, device=DEV)
    outs = {}
    for lazy in (False, True):
        hm = HookManager(keys=['g'])
        hm.register('g', RecencyNeighborHook(num_nodes=N, num_nbrs=[k, k],
                                             seed_nodes_keys=['edge_src', 'edge_dst'],
                                             seed_times_keys=['edge_time', 'edge_time'],
                                             lazy_edge_x=lazy))
        res = []
        with hm.activate('g'), torch.no_grad():
            for b, batch in enumerate(DGDataLoader(dg, batch_size=bs, hook_manager=hm)):
                if b % 9 == 0 or b == E // bs - 1:
                    assert isinstance(batch.nbr_edge_x[1], LazyEdgeRows) == lazy
                    res.append(model(node_x, batch.seed_nids, batch.seed_times, batch.nbr_nids,
                                     batch.nbr_edge_x, batch.nbr_edge_time))
        outs[lazy] = res
    assert len(outs[True]) == len(outs[False]) > 3
    for a, b in zip(outs[True], outs[False]):
        assert torch.equal(a, b)


def test_lazy_rows_behave_as_tensors():
    table = torch.arange(24., device=DEV).reshape(6, 4)
    rows = torch.tensor([[0, -1, 5], [2, 2, -1]], dtype=torch.int32, device=DEV)
    lz = LazyEdgeRows(table, rows)
    dense = lz.materialize()
    assert lz.shape == (2, 3, 4) and lz.dtype == torch.float32 and lz.is_cuda
    assert torch.equal(dense[0, 1], torch.zeros(4, device=DEV)) and torch.equal(dense[1, 0], table[2])
    assert torch.equal(torch.cat([lz, lz]), torch.cat([dense, dense]))
    assert torch.equal(lz.flatten(0, -2), dense.flatten(0, -2)) and torch.equal(lz * 2, dense * 2)
    assert torch.equal(lz[1], dense[1]) and isinstance(lz[0:1], LazyEdgeRows)
    assert [v.shape for v in lz.split(1)] == [(1, 3, 4), (1, 3, 4)]
