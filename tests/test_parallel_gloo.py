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
