"""Where the 12 us per batch of the drop-in loader path go (host side)."""
import cProfile, pstats, sys, time, torch
sys.path.insert(0, '.')
from tgm_b200 import DGData, DGDataLoader, DGraph, HookManager, RecencyNeighborHook, RandomNegativeEdgeSamplerHook
dev = torch.device('cuda', 0)
E, N, D, bs, k = 4_000_000, 1_000_000, 16, 200, 20
g = torch.Generator().manual_seed(0)
src = torch.randint(0, N, (E,), generator=g, dtype=torch.int32)
dst = torch.randint(0, N, (E,), generator=g, dtype=torch.int32)
t = torch.sort(torch.randint(0, 2000, (E,), generator=g))[0]
x = torch.randn(E, D, generator=g)
dg = DGraph(DGData.from_raw(t, torch.stack([src, dst], 1), x), device=dev)
for with_neg in (False, True):
    hm = HookManager(keys=['b'])
    keys_n, keys_t = ['edge_src', 'edge_dst'], ['edge_time', 'edge_time']
    if with_neg:
        hm.register('b', RandomNegativeEdgeSamplerHook(low=0, high=N))
        keys_n, keys_t = keys_n + ['neg'], keys_t + ['neg_time']
    hm.register('b', RecencyNeighborHook(num_nodes=N, num_nbrs=[k], seed_nodes_keys=keys_n, seed_times_keys=keys_t))
    with hm.activate('b'):
        for p in range(2):
            hm.reset_state(); torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
            for batch in DGDataLoader(dg, batch_size=bs, hook_manager=hm):
                n += 1
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f'with_neg={with_neg}: {dt / n * 1e6:.2f} us/batch over {n} batches')
        hm.reset_state()
        pr = cProfile.Profile(); pr.enable()
        for batch in DGDataLoader(dg, batch_size=bs, hook_manager=hm):
            pass
        pr.disable()
        pstats.Stats(pr).sort_stats('tottime').print_stats(14)
