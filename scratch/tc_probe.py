import sys, torch
sys.path.insert(0, '.')
from tgm_b200 import _cabi
dev = 'cuda:0'
S, N, K = 25600, 600, 200
A = torch.randn(S, K, device=dev); W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
out = torch.empty(S, N, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
for _ in range(3):
    _cabi.check(_cabi.lib.tgm_tc_linear(S, N, K, A.data_ptr(), W.data_ptr(), b.data_ptr(), None, 0, out.data_ptr(), st))
torch.cuda.synchronize()
