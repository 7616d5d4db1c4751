"""Host-link probe: D2H / H2D bandwidth per rank with all ranks copying at once, for different
host allocations (torch pinned = cudaHostAlloc; anonymous mmap + MADV_HUGEPAGE + cudaHostRegister).
Run alone or under torchrun."""
import ctypes, mmap, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); lr = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr); dev = f'cuda:{lr}'
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device(dev))

def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

def rmax(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

NB = 128 << 20
if rank == 0:
    for f in ('/sys/kernel/mm/transparent_hugepage/enabled', '/sys/kernel/mm/transparent_hugepage/defrag', '/proc/sys/vm/nr_hugepages'):
        try: print(f, open(f).read().strip(), flush=True)
        except Exception as e: print(f, e)
    print('cpus', len(os.sched_getaffinity(0)), 'numa nodes', [d for d in os.listdir('/sys/devices/system/node') if d.startswith('node')], flush=True)

def thp_registered(nbytes):
    mm = mmap.mmap(-1, nbytes + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    try: mm.madvise(mmap.MADV_HUGEPAGE)
    except Exception as e: print('madvise', e)
    a = np.frombuffer(mm, dtype=np.uint8)
    off = (-a.ctypes.data) % (2 << 20)
    a = a[off:off + nbytes]
    a[:] = 1  # touch
    t = torch.from_numpy(a)
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), nbytes, 0)
    assert int(rc) == 0, rc
    return t, mm

modes = {'cudaHostAlloc (torch pinned)': lambda: (torch.empty(NB, dtype=torch.uint8, pin_memory=True), None),
         'mmap + MADV_HUGEPAGE + cudaHostRegister': lambda: thp_registered(NB)}
d = torch.empty(NB, dtype=torch.uint8, device=dev)
for name, mk in modes.items():
    h, keep = mk()
    res = {}
    for dirn, (dst_, src_) in (('d2h', (h, d)), ('h2d', (d, h))):
        dst_.copy_(src_, non_blocking=True)
        barrier(); t0 = time.perf_counter()
        for _ in range(10):
            dst_.copy_(src_, non_blocking=True)
        barrier()
        res[dirn] = 10 * NB / rmax(time.perf_counter() - t0) / 1e9
    if rank == 0:
        print(f'world={world} {name}: d2h {res["d2h"]:.1f} GB/s per rank, h2d {res["h2d"]:.1f} GB/s per rank', flush=True)
    if keep is not None:
        torch.cuda.cudart().cudaHostUnregister(h.data_ptr())
    del h

def measure(h, d, n_copies, label, both=False):
    res = {}
    for dirn, (dst_, src_) in (('d2h', (h, d)), ('h2d', (d, h))):
        dst_.copy_(src_, non_blocking=True)
        barrier(); t0 = time.perf_counter()
        for _ in range(n_copies):
            dst_.copy_(src_, non_blocking=True)
        barrier()
        mine = n_copies * h.numel() / (time.perf_counter() - t0) / 1e9
        res[dirn] = (n_copies * h.numel() / rmax(time.perf_counter() - t0) / 1e9, mine)
    allr = [None] * world
    if world > 1:
        dist.all_gather_object(allr, (res['d2h'][1], res['h2d'][1]))
    else:
        allr = [(res['d2h'][1], res['h2d'][1])]
    if rank == 0:
        print(f'world={world} {label}: d2h {res["d2h"][0]:.1f} h2d {res["h2d"][0]:.1f} GB/s per rank (slowest); per-rank d2h '
              + ' '.join(f'{a:.0f}' for a, _ in allr), flush=True)

h128 = torch.empty(128 << 20, dtype=torch.uint8, pin_memory=True)
measure(h128, d, 10, '128 MB x10, pin_memory=True')
h64 = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
d64 = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
measure(h64, d64, 8, '64 MB x8, empty().pin_memory()  [bench.py link]')
big = torch.empty(24 << 30, dtype=torch.uint8, device=dev)
hostbig = [torch.empty(256 << 20, dtype=torch.uint8).pin_memory() for _ in range(4)]
measure(h64, d64, 8, '64 MB x8 after 24 GB device + 1 GB pinned allocations')
measure(h128, d, 10, '128 MB x10 again')
# concurrent: all ranks D2H on one stream while H2D runs on another
s2 = torch.cuda.Stream(dev)
hin = torch.empty(4 << 20, dtype=torch.uint8, pin_memory=True); din = torch.empty(4 << 20, dtype=torch.uint8, device=dev)
barrier(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s2):
        din.copy_(hin, non_blocking=True)
    h128.copy_(d, non_blocking=True)
barrier()
v = 10 * h128.numel() / rmax(time.perf_counter() - t0) / 1e9
if rank == 0:
    print(f'world={world} 128 MB D2H x10 with 4 MB H2D on a second stream: {v:.1f} GB/s per rank', flush=True)
if world > 1:
    dist.destroy_process_group()
