"""Launches of tc3_linear_kernel / the CUTLASS collective on DyGFormer's shapes at 12800 tokens (for ncu)."""
import sys, torch
sys.path.insert(0, '.')
from tgm_b200 import _cabi
dev = 'cuda:0'; st = torch.cuda.current_stream(dev).cuda_stream
for (S, N, K, g_, r_) in [(12800, 600, 200, 0, 0), (12800, 800, 200, 1, 0), (12800, 200, 800, 0, 1)]:
    A = torch.randn(S, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    R = torch.randn(S, N, device=dev); out = torch.empty(S, N, device=dev)
    for _ in range(3):
        _cabi.check(_cabi.lib.tgm_tc_linear(S, N, K, A.data_ptr(), W.data_ptr(), b.data_ptr(), R.data_ptr() if r_ else None, g_, out.data_ptr(), st))
        _cabi.check(_cabi.lib.tgm_fastf32_linear(S, N, K, A.data_ptr(), W.data_ptr(), b.data_ptr(), R.data_ptr() if r_ else None, g_, out.data_ptr(), st))
    torch.cuda.synchronize()
