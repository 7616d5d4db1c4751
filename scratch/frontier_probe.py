"""Correctness + timing of tgm_frontier_compact over sizes (incl. the multi-launch range)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgm_b200 import _cabi

DEV = 'cuda:0'
st = torch.cuda.current_stream(DEV).cuda_stream


def run(nid, idx, cnt):
    _cabi.check(_cabi.lib.tgm_frontier_compact(nid.data_ptr(), nid.numel(), idx.data_ptr(), cnt.data_ptr(), st))


for n, p in [(1, 0.5), (31, 0.3), (4096, 0.37), (4097, 0.37), (1_000_003, 0.37), (4_000_000, 1.0), (16_000_000, 1.0), (40_000_000, 0.001),
             (40_000_000, 0.5), (40_000_000, 1.0), (260_000_000, 0.3)]:
    g = torch.Generator(device=DEV).manual_seed(n)
    nid = torch.randint(0, 1000, (n,), generator=g, device=DEV, dtype=torch.int32)
    pad = torch.rand(n, generator=g, device=DEV) < p
    nid[pad] = -1
    del pad
    idx = torch.empty(n, dtype=torch.int64, device=DEV)
    cnt = torch.full((1,), -7, dtype=torch.int64, device=DEV)
    run(nid, idx, cnt)
    want = torch.nonzero(nid != -1).reshape(-1)
    c = int(cnt.item())
    ok = c == want.numel() and torch.equal(idx[:c], want)
    kept = want.numel()
    del want
    for _ in range(3):
        run(nid, idx, cnt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        run(nid, idx, cnt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = (4 * n + 8 * kept) / 1e9
    print(f'n={n} pad={p}: ok={ok} kept={kept} {ms*1e3:.1f} us  {gb/ms*1e3:.0f} GB/s algorithmic ({gb/ms*1e3/6552.3:.2f} of peak)', flush=True)
    del nid, idx
