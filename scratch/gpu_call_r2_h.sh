#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_linear.py -q > gpurun_out/h_tc.log 2>&1; tail -12 gpurun_out/h_tc.log
timeout 600 python -m pytest tests/test_gpu_nn.py tests/test_gpu_lazy_edge_x.py -q > gpurun_out/h_nn.log 2>&1; tail -25 gpurun_out/h_nn.log
timeout 300 python bench_rows.py --rows dygformer 2>&1 | cut -c1-300
python - <<'PY'
import torch, ctypes, sys
sys.path.insert(0,'.')
from tgm_b200 import _cabi
dev='cuda:0'
for (S,N,K,g_,r_) in [(25600,600,200,0,0),(25600,200,200,0,1),(25600,800,200,1,0),(25600,200,800,0,1)]:
    A=torch.randn(S,K,device=dev); W=torch.randn(N,K,device=dev); b=torch.randn(N,device=dev); R=torch.randn(S,N,device=dev); out=torch.empty(S,N,device=dev)
    st=torch.cuda.current_stream(dev).cuda_stream
    f=lambda: _cabi.check(_cabi.lib.tgm_tc_linear(S,N,K,A.data_ptr(),W.data_ptr(),b.data_ptr(),R.data_ptr() if r_ else None,g_,out.data_ptr(),st))
    for _ in range(3): f()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/20
    g2=lambda: torch.nn.functional.linear(A,W,b)
    for _ in range(3): g2()
    e0.record()
    for _ in range(20): g2()
    e1.record(); torch.cuda.synchronize(); ms2=e0.elapsed_time(e1)/20
    print(f'tc_linear {S}x{N}x{K}: {ms*1e3:.1f} us  ({2*S*N*K/ms/1e9:.1f} TFLOP/s fp32-equivalent); torch fp32 {ms2*1e3:.1f} us')
PY
