import sys, time, torch
sys.path.insert(0, '.')
t0 = time.perf_counter()
from tgm_b200 import _cabi
from tgm_b200.core.storage import DeviceCOOStorage
from tgm_b200.sampler import RecencyCSR
print(f'import {time.perf_counter() - t0:.3f} s')
dev = torch.device('cuda', 0)
E, N, D, bs = 100_000_000, 1_000_000, 16, 200
gen = torch.Generator(device=dev).manual_seed(0)
src = torch.randint(0, N, (E,), generator=gen, device=dev, dtype=torch.int32)
dst = torch.randint(0, N, (E,), generator=gen, device=dev, dtype=torch.int32)
t = torch.sort(torch.randint(0, 2000, (E,), generator=gen, device=dev))[0]
x = torch.randn((E, D), generator=gen, device=dev)
torch.cuda.synchronize()
_cabi.check(_cabi.lib.tgm_set_option(b'trace', 1))
for i in range(3):
    t0 = time.perf_counter()
    store = DeviceCOOStorage.from_device_tensors(src, dst, t, x, N)
    t1 = time.perf_counter()
    csr = RecencyCSR(store, bs, colocate_x=True)
    torch.cuda.synchronize()
    print(f'build {i}: store {t1 - t0:.3f} s, adjacency {time.perf_counter() - t1:.3f} s', flush=True)
    del csr, store
