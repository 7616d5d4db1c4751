import sys, torch
sys.path.insert(0, '.')
from tgm_b200 import _cabi
dev = 'cuda:0'
S, N, K = 25600, 600, 200
A = torch.randn(S, K, device=dev); W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
out = torch.empty(S, N, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
def run(): _cabi.check(_cabi.lib.tgm_tc_linear(S, N, K, A.data_ptr(), W.data_ptr(), b.data_ptr(), None, 0, out.data_ptr(), st))
for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
    _cabi.check(_cabi.lib.tgm_set_option(b'tc_debug', dbg))
    for _ in range(3): run()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    print(f'dbg={dbg} (no: {"MMA " if dbg&1 else ""}{"drain " if dbg&2 else ""}{"staging" if dbg&4 else ""}): {e0.elapsed_time(e1)/20*1e3:.1f} us')
