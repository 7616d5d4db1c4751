"""tc3_linear column-tile width sweep on DyGFormer's four token-linear shapes."""
import sys, torch
sys.path.insert(0, '.')
from tgm_b200 import _cabi
dev = 'cuda:0'; st = torch.cuda.current_stream(dev).cuda_stream
SHAPES = [(12800, 600, 200, 0, 0), (12800, 200, 200, 0, 1), (12800, 800, 200, 1, 0), (12800, 200, 800, 0, 1)]
for (S, N, K, g_, r_) in SHAPES:
    A = torch.randn(S, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    R = torch.randn(S, N, device=dev); out = torch.empty(S, N, device=dev)
    want = (A.double() @ W.double().T + b.double() + (R.double() if r_ else 0))
    if g_ == 1: want = torch.nn.functional.gelu(want)
    f = lambda: _cabi.check(_cabi.lib.tgm_tc_linear(S, N, K, A.data_ptr(), W.data_ptr(), b.data_ptr(), R.data_ptr() if r_ else None, g_, out.data_ptr(), st))
    line = []
    for bn in (0, 200, 160, 152, 136, 120, 104, 88, 72):
        _cabi.check(_cabi.lib.tgm_set_option(b'tc_bn', bn))
        for _ in range(3): f()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
        err = (out.double() - want).abs().max().item()
        line.append(f'bn={bn}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us (err {err:.1e})')
    if g_ != 2:
        ff = lambda: _cabi.check(_cabi.lib.tgm_fastf32_linear(S, N, K, A.data_ptr(), W.data_ptr(), b.data_ptr(), R.data_ptr() if r_ else None, g_, out.data_ptr(), st))
        for _ in range(3): ff()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        for _ in range(20): ff()
        e1.record(); torch.cuda.synchronize()
        line.append(f'CUTLASS FastF32: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us')
    print(f'{S}x{N}x{K}: ' + ' | '.join(line), flush=True)
_cabi.check(_cabi.lib.tgm_set_option(b'tc_bn', 0))
