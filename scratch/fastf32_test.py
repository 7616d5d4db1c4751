import ctypes, os, torch, time
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fastf32_try.so'))
lib.fastf32_gemm_nt.argtypes = [ctypes.c_int]*3 + [ctypes.c_void_p]*4 + [ctypes.c_size_t, ctypes.c_void_p]
lib.fastf32_workspace.restype = ctypes.c_size_t
torch.backends.cuda.matmul.allow_tf32 = False
dev = 'cuda:0'
for (M, N, K) in [(25600, 800, 200), (25600, 200, 800), (25600, 600, 200), (25600, 200, 200), (400, 172, 200), (12000, 204, 273 - 1), (1000, 100, 100)]:
    g = torch.Generator(device=dev).manual_seed(0)
    A = torch.randn(M, K, device=dev, generator=g); W = torch.randn(N, K, device=dev, generator=g) * 0.1
    C = torch.empty(M, N, device=dev)
    ws_bytes = lib.fastf32_workspace(M, N, K)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.fastf32_gemm_nt(M, N, K, A.data_ptr(), W.data_ptr(), C.data_ptr(), ws.data_ptr(), ws_bytes, st)
    torch.cuda.synchronize()
    if rc != 0:
        print((M, N, K), 'rc', rc); continue
    ref64 = (A.double() @ W.double().T)
    ref32 = A @ W.T
    e_fast = (C.double() - ref64).abs().max().item(); e_32 = (ref32.double() - ref64).abs().max().item()
    def tm(fn, it=20):
        for _ in range(3): fn()
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(it): fn()
        b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it
    t_fast = tm(lambda: lib.fastf32_gemm_nt(M, N, K, A.data_ptr(), W.data_ptr(), C.data_ptr(), ws.data_ptr(), ws_bytes, st))
    t_32 = tm(lambda: torch.matmul(A, W.T, out=ref32))
    print((M, N, K), f'ws {ws_bytes} err fast {e_fast:.3e} err fp32 {e_32:.3e} max|ref| {ref64.abs().max().item():.2f}  t_fast {t_fast*1e3:.1f} us ({2*M*N*K/t_fast/1e9:.1f} TF)  t_cublas_fp32 {t_32*1e3:.1f} us ({2*M*N*K/t_32/1e9:.1f} TF)')
