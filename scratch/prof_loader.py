import cProfile, pstats, sys, os, io
sys.path.insert(0, os.getcwd())
import torch
from tgm_b200 import DGDataLoader, DGraph, HookManager, RecencyNeighborHook
from tgm_b200.core.storage import DeviceCOOStorage, DGSliceTracker
from tgm_b200.core.timedelta import TimeDeltaDG
dev = torch.device('cuda', 0)
E, N, D, k, bs = 4_000_000, 1_000_000, 16, 20, 200
g = torch.Generator(device=dev).manual_seed(0)
src = torch.randint(0, N, (E,), generator=g, device=dev, dtype=torch.int32)
dst = torch.randint(0, N, (E,), generator=g, device=dev, dtype=torch.int32)
t = torch.sort(torch.randint(0, 2000, (E,), generator=g, device=dev))[0]
x = torch.randn((E, D), generator=g, device=dev)
store = DeviceCOOStorage.from_device_tensors(src, dst, t, x, N)
dg = DGraph._from_storage(store, TimeDeltaDG('r'), dev, DGSliceTracker(end_idx=E))
from tgm_b200 import RandomNegativeEdgeSamplerHook
for wb in ('default', 'default+neg', 0):
    hm = HookManager(keys=['bench'])
    kw = {} if wb != 0 else {'window_batches': 0}
    if wb == 'default+neg':
        hm.register('bench', RandomNegativeEdgeSamplerHook(low=0, high=N))
        hm.register('bench', RecencyNeighborHook(num_nodes=N, num_nbrs=[k], seed_nodes_keys=['edge_src', 'edge_dst', 'neg'],
                    seed_times_keys=['edge_time', 'edge_time', 'neg_time']))
    else:
        hm.register('bench', RecencyNeighborHook(num_nodes=N, num_nbrs=[k], seed_nodes_keys=['edge_src', 'edge_dst'],
                    seed_times_keys=['edge_time', 'edge_time'], **kw))
    def run():
        got = 0
        with hm.activate('bench'):
            hm.reset_state()
            for batch in DGDataLoader(dg, batch_size=bs, hook_manager=hm):
                got += batch.nbr_nids[0].numel()
        torch.cuda.synchronize()
        return got
    run()
    import time
    t0 = time.perf_counter(); run(); dt = time.perf_counter() - t0
    print(f'window_batches={wb}: {dt / (E // bs) * 1e6:.1f} us/batch')
    pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:6000])
