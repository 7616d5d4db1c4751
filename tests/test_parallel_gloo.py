"""N>1 host logic on CPU: time-range sharding of the batch stream, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tgm_b200.parallel import (current_shard, gather_shard_summaries, max_over_ranks,
                               shard_batches, sum_over_ranks)


@pytest.mark.parametrize('E,bs,world', [(1000, 200, 1), (1001, 200, 2), (157474, 200, 8),
                                        (5, 200, 4), (10_000_000, 200, 8), (999, 7, 3)])
def test_shards_partition_the_batch_stream(E, bs, world):
    shards = [shard_batches(E, bs, r, world) for r in range(world)]
    assert shards[0].edge_lo == 0 and shards[-1].edge_hi == E
    for a, b in zip(shards, shards[1:]):
        assert a.edge_hi == b.edge_lo and a.batch_hi == b.batch_lo
    sizes = [s.num_batches for s in shards]
    assert max(sizes) - min(sizes) <= 1
    for s in shards:
        assert s.edge_lo % bs == 0  # shard (and window) starts sit on loader-batch boundaries
        wins = list(s.windows(50, bs))
        assert sum(hi - lo for lo, hi in wins) == s.num_edges
        assert all(lo % bs == 0 for lo, _ in wins)


def test_shard_argument_errors():
    with pytest.raises(ValueError):
        shard_batches(10, 2, 2, 2)
    with pytest.raises(ValueError):
        shard_batches(10, 0, 0, 1)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, E: int, bs: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sh = current_shard(E, bs)
        # every rank "samples" its shard: here the per-shard summary is (edges, seeds, a checksum
        # of the stream positions it owns); the gather must reassemble the whole stream exactly
        pos = np.arange(sh.edge_lo, sh.edge_hi, dtype=np.int64)
        summ = gather_shard_summaries([sh.num_edges, 2 * sh.num_edges, float(pos.sum())])
        total_edges = sum_over_ranks(sh.num_edges)
        slowest = max_over_ranks(float(rank + 1))
        if rank == 0:
            np.save(os.path.join(out_dir, 'summ.npy'), summ.numpy())
            np.save(os.path.join(out_dir, 'misc.npy'), np.array([total_edges, slowest]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    E, bs, world = 100_123, 200, 2
    mp.spawn(_worker, args=(world, _free_port(), E, bs, str(tmp_path)), nprocs=world, join=True)
    summ = np.load(tmp_path / 'summ.npy')
    total_edges, slowest = np.load(tmp_path / 'misc.npy')
    assert summ.shape == (2, 3)
    assert summ[:, 0].sum() == E == total_edges
    assert summ[:, 2].sum() == E * (E - 1) / 2  # every stream position owned exactly once
    assert slowest == 2.0


def _merge_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    from tgm_b200.parallel import merge_node_memory
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        N, M = 50, 4
        g = torch.Generator().manual_seed(0)
        base_mem = torch.randn(N, M, generator=g)           # common snapshot
        base_lu = torch.randint(0, 100, (N,), generator=g)
        # rank r touches nodes with (i % 3 == r) or (i % 7 == 0); its rows get a rank signature
        idx = torch.arange(N)
        touched = (idx % 3 == rank) | (idx % 7 == 0)
        mem, lu = base_mem.clone(), base_lu.clone()
        mem[touched] = 1000.0 * (rank + 1) + idx[touched, None].float()
        lu[touched] = 1000 * (rank + 1) + idx[touched]
        merge_node_memory(mem, lu, touched)
        if rank == 0:
            torch.save({'mem': mem, 'lu': lu, 'base_mem': base_mem, 'base_lu': base_lu},
                       os.path.join(out_dir, 'merged.pt'))
        # every rank must hold the same merged state
        ref = [torch.empty_like(mem) for _ in range(world)]
        dist.all_gather(ref, mem)
        assert all(torch.equal(r, mem) for r in ref)
    finally:
        dist.destroy_process_group()


def test_merge_node_memory_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_merge_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    z = torch.load(tmp_path / 'merged.pt')
    idx = torch.arange(50)
    for i in idx.tolist():
        owners = [r for r in range(world) if (i % 3 == r) or (i % 7 == 0)]
        if owners:
            r = max(owners)  # the later shard wins
            assert z['mem'][i, 0] == 1000.0 * (r + 1) + i and z['lu'][i] == 1000 * (r + 1) + i
        else:
            assert torch.equal(z['mem'][i], z['base_mem'][i]) and z['lu'][i] == z['base_lu'][i]


def _join_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    """join_node_memory (touched rows packed, ONE all-gather, scatter in rank order) must leave
    what the dense merge leaves; rank `world - 1` touches nothing when world == 3."""
    from tgm_b200.parallel import join_node_memory, merge_node_memory
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        N, M = 61, 8
        g = torch.Generator().manual_seed(0)
        base_mem = torch.randn(N, M, generator=g)
        base_lu = torch.randint(0, 100, (N,), generator=g)
        idx = torch.arange(N)
        touched = (idx % 3 == rank) | (idx % 7 == 0) | ((idx % 11 == 0) & (rank == 1))
        if world == 3 and rank == 2:
            touched = torch.zeros(N, dtype=torch.bool)
        mem, lu = base_mem.clone(), base_lu.clone()
        mem[touched] = 1000.0 * (rank + 1) + idx[touched, None].float() + torch.arange(M).float()
        lu[touched] = 1000 * (rank + 1) + idx[touched]
        dense_mem, dense_lu = mem.clone(), lu.clone()
        merge_node_memory(dense_mem, dense_lu, touched)
        info = join_node_memory(mem, lu, idx[touched].to(torch.int32))
        assert torch.equal(mem, dense_mem) and torch.equal(lu, dense_lu)
        assert info['counts'][rank] == int(touched.sum()) and info['row_bytes'] == 16 + 4 * M
        assert info['recv_bytes'] == sum(info['counts']) * info['row_bytes']
        ref = [torch.empty_like(mem) for _ in range(world)]
        dist.all_gather(ref, mem)
        assert all(torch.equal(r, mem) for r in ref)
        if rank == 0:
            torch.save({'mem': mem, 'lu': lu, 'base_mem': base_mem, 'base_lu': base_lu,
                        'counts': info['counts']}, os.path.join(out_dir, 'joined.pt'))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_join_node_memory_gloo(tmp_path, world):
    mp.spawn(_join_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    z = torch.load(tmp_path / 'joined.pt')
    for i in range(61):
        owners = [r for r in range(world) if not (world == 3 and r == 2) and
                  ((i % 3 == r) or (i % 7 == 0) or (i % 11 == 0 and r == 1))]
        if owners:
            r = max(owners)  # the later time shard wins
            assert z['mem'][i, 0] == 1000.0 * (r + 1) + i and z['lu'][i] == 1000 * (r + 1) + i
        else:
            assert torch.equal(z['mem'][i], z['base_mem'][i]) and z['lu'][i] == z['base_lu'][i]
    assert len(z['counts']) == world and (world == 2 or z['counts'][2] == 0)


def test_join_node_memory_is_a_no_op_outside_torch_distributed():
    from tgm_b200.parallel import join_node_memory
    mem, lu = torch.randn(5, 4), torch.arange(5)
    want = mem.clone()
    info = join_node_memory(mem, lu, torch.tensor([1, 3], dtype=torch.int32))
    assert torch.equal(mem, want) and info['counts'] == [2] and info['recv_bytes'] == 0


def _grad_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    from tgm_b200.parallel import average_gradients
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        frozen = torch.nn.Parameter(torch.ones(2), requires_grad=False)
        unused = torch.nn.Parameter(torch.ones(4))  # no gradient on rank 1
        x = torch.full((2, 5), float(rank + 1))
        loss = lin(x).sum() + (unused.sum() * 3 if rank == 0 else 0)
        loss.backward()
        average_gradients([lin.weight, lin.bias, frozen, unused])
        torch.save({'w': lin.weight.grad, 'b': lin.bias.grad, 'u': unused.grad, 'f': frozen.grad},
                   os.path.join(out_dir, f'grads{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_average_gradients_world_size_2_gloo(tmp_path):
    """Sharded training: every rank ends with the mean of the per-rank gradients, including a
    parameter that received no gradient on one rank; frozen parameters are left alone."""
    world = 2
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g0, g1 = (torch.load(tmp_path / f'grads{r}.pt') for r in range(world))
    for k in ('w', 'b', 'u'):
        assert torch.equal(g0[k], g1[k]), k
    # d/dW sum(W x + b) = sum over the 2 rows of x: 2*(rank+1) per entry -> mean over ranks = 3
    assert torch.allclose(g0['w'], torch.full((3, 5), 3.0)) and torch.allclose(g0['b'], torch.full((3,), 2.0))
    assert torch.allclose(g0['u'], torch.full((4,), 1.5)) and g0['f'] is None
