import ctypes, os, torch
here = os.path.dirname(os.path.abspath(__file__))
torch.backends.cuda.matmul.allow_tf32 = False
dev = 'cuda:0'
def tm(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it
shapes = [(25600, 800, 200), (25600, 200, 800), (25600, 600, 200), (25600, 200, 200), (12000, 204, 272), (600, 344, 444)]
for v in ['v1sm_tmem', 'v2sm_smem', 'v2sm_tmem', 'v1sm_tmem64']:
    lib = ctypes.CDLL(os.path.join(here, f'ff_{v}.so'))
    fn = getattr(lib, f'{v}_gemm')
    fn.argtypes = [ctypes.c_int]*3 + [ctypes.c_void_p]*4 + [ctypes.c_size_t, ctypes.c_void_p]
    ws = torch.empty(1 << 24, dtype=torch.uint8, device=dev)
    for (M, N, K) in shapes:
        g = torch.Generator(device=dev).manual_seed(0)
        A = torch.randn(M, K, device=dev, generator=g); W = torch.randn(N, K, device=dev, generator=g) * 0.1
        C = torch.zeros(M, N, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        call = lambda: fn(M, N, K, A.data_ptr(), W.data_ptr(), C.data_ptr(), ws.data_ptr(), ws.numel(), st)
        rc = call(); torch.cuda.synchronize()
        if rc: print(v, (M, N, K), 'rc', rc); continue
        err = (C.double() - A.double() @ W.double().T).abs().max().item()
        t = tm(call)
        print(v, (M, N, K), f'err {err:.2e}  {t*1e3:.1f} us  {2*M*N*K/t/1e9:.1f} TF')
