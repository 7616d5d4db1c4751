import sys, torch
sys.path.insert(0, '.')
from tgm_b200 import _cabi
dev = 'cuda:0'; st = torch.cuda.current_stream(dev).cuda_stream
tot = [0, 0, 0]
SHAPES = [(25600, 600, 200, 0, 0), (25600, 200, 200, 0, 1), (25600, 800, 200, 1, 0), (25600, 200, 800, 0, 1)]
TGAT = [(12600, 104, 548, 0, 0), (12600, 172, 104, 2, 0), (12600, 172, 172, 0, 0), (600, 104, 548, 0, 0), (600, 172, 104, 2, 0), (600, 172, 172, 0, 0), (600, 272, 888, 0, 0), (600, 172, 444, 2, 0)]
for (S, N, K, g_, r_) in SHAPES + TGAT:
    A = torch.randn(S, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    R = torch.randn(S, N, device=dev); out = torch.empty(S, N, device=dev)
    fns = [lambda: _cabi.check(_cabi.lib.tgm_tc_linear(S, N, K, A.data_ptr(), W.data_ptr(), b.data_ptr(), R.data_ptr() if r_ else None, g_, out.data_ptr(), st)),
           (lambda: _cabi.check(_cabi.lib.tgm_fastf32_linear(S, N, K, A.data_ptr(), W.data_ptr(), b.data_ptr(), R.data_ptr() if r_ else None, g_, out.data_ptr(), st))) if g_ != 2 else (lambda: None),
           lambda: torch.nn.functional.linear(A, W, b)]
    us = []
    for f in fns:
        for _ in range(3): f()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize(); us.append(e0.elapsed_time(e1) / 20 * 1e3)
    if S == 25600:
        for i in range(3): tot[i] += us[i]
    print(f'{S}x{N}x{K}: tc_linear {us[0]:.1f} us ({2*S*N*K/us[0]/1e6:.1f} TFLOP/s-equiv) | CUTLASS FastF32 {us[1]:.1f} us | torch fp32 {us[2]:.1f} us')
print(f'sum of the four DyGFormer linears: tc_linear {tot[0]:.0f} us | CUTLASS FastF32 {tot[1]:.0f} us | torch fp32 {tot[2]:.0f} us')
