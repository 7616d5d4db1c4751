#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/p_gpu_tests.log 2>&1; tail -25 gpurun_out/p_gpu_tests.log
python - <<'PY'
# time-unit loader throughput: default (windowed) vs ring
import sys, time, numpy as np, torch
sys.path.insert(0,'.')
from tgm_b200 import DGData, DGDataLoader, DGraph, HookManager, RecencyNeighborHook
rng=np.random.default_rng(0)
E,N,D=1_000_000,100_000,16
src=rng.integers(0,N,E).astype(np.int32); dst=rng.integers(0,N,E).astype(np.int32)
t=np.sort(rng.integers(0,5000*200,E)).astype(np.int64); x=rng.standard_normal((E,D)).astype(np.float32)
dg=DGraph(DGData.from_raw(torch.from_numpy(t), torch.from_numpy(np.stack([src,dst],1)), torch.from_numpy(x), time_delta='s'), device='cuda:0')
for wb in (None, 0):
    kw={} if wb is None else {'window_batches':wb}
    hm=HookManager(keys=['g']); hm.register('g', RecencyNeighborHook(num_nodes=N,num_nbrs=[20],seed_nodes_keys=['edge_src','edge_dst'],seed_times_keys=['edge_time','edge_time'],**kw))
    with hm.activate('g'):
        for rep in range(2):
            hm.reset_state(); torch.cuda.synchronize(); t0=time.perf_counter(); nb=0
            for b in DGDataLoader(dg,batch_size=200,batch_unit='s',hook_manager=hm): nb+=1
            torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(f'time-unit batches (200 s windows, ~200 edges each), window={wb}: {dt/nb*1e6:.1f} us/batch over {nb} batches')
PY
