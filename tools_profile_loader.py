import cProfile, pstats, time, sys, numpy as np, torch
sys.path.insert(0, '.')
from bench_rows import wiki_stream, graph
from tgm_b200 import DGDataLoader, HookManager, RecencyNeighborHook
src, dst, t, x, N = wiki_stream()
dg = graph(src, dst, t, x)
for W in (0, 200):
    hm = HookManager(keys=['g'])
    hm.register('g', RecencyNeighborHook(num_nodes=N, num_nbrs=[10], seed_nodes_keys=['edge_src', 'edge_dst'],
                                         seed_times_keys=['edge_time', 'edge_time'], window_batches=W))
    with hm.activate('g'):
        for _ in DGDataLoader(dg, batch_size=200, hook_manager=hm): pass
        hm.reset_state(); torch.cuda.synchronize()
        t0 = time.perf_counter(); nb = 0
        for _ in DGDataLoader(dg, batch_size=200, hook_manager=hm): nb += 1
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f'window={W}: {dt/nb*1e6:.1f} us/batch, {2*len(src)*10/dt/1e6:.1f} M sampled-edges/s')
        hm.reset_state()
        pr = cProfile.Profile(); pr.enable()
        for _ in DGDataLoader(dg, batch_size=200, hook_manager=hm): pass
        pr.disable()
        pstats.Stats(pr).sort_stats('tottime').print_stats(12)
