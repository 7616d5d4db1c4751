"""Where the first epochs of the default hook spend their time (wiki-shaped row of bench_rows.py)."""
import sys, time
sys.path.insert(0, '.')
import torch
from bench_rows import wiki_stream, graph, DEV
from tgm_b200 import DGDataLoader, HookManager, RecencyNeighborHook
src, dst, t, x, N = wiki_stream()
dg = graph(src, dst, t, x)
bs, k = 200, 10
hm = HookManager(keys=['g'])
hm.register('g', RecencyNeighborHook(num_nodes=N, num_nbrs=[k], seed_nodes_keys=['edge_src', 'edge_dst'], seed_times_keys=['edge_time', 'edge_time']))
with hm.activate('g'):
    for ep in range(7):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hm.reset_state()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        it = iter(DGDataLoader(dg, batch_size=bs, hook_manager=hm))
        t2 = time.perf_counter()
        b = next(it)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        n = 1
        for b in it:
            n += 1
        t4 = time.perf_counter()
        torch.cuda.synchronize(); t5 = time.perf_counter()
        print(f'epoch {ep}: reset {1e3*(t1-t0):.2f} ms | loader ctor {1e3*(t2-t1):.2f} | first batch (window) {1e3*(t3-t2):.2f} | other {n-1} batches {1e3*(t4-t3):.2f} ({1e6*(t4-t3)/(n-1):.1f} us each) | final sync {1e3*(t5-t4):.2f} | reserved {torch.cuda.memory_reserved()/2**30:.2f} GiB', flush=True)
