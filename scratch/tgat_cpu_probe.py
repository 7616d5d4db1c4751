"""Is TGAT.forward bound by the host or by the device?  Wall time of N un-synchronised forwards
(host issue rate) vs the device time of the same N (CUDA events)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
sys.argv = [sys.argv[0], '0']
exec(open('scratch/tgat_probe.py').read().split('with torch.no_grad():')[0])
N_IT = 200
with torch.no_grad():
    for _ in range(10): model(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(N_IT): model(*args)
    t_issue = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
print(f'host issue {t_issue / N_IT * 1e6:.0f} us/forward | device {e0.elapsed_time(e1) / N_IT * 1e3:.0f} us/forward | wall {t_wall / N_IT * 1e6:.0f} us')
import cProfile, pstats
pr = cProfile.Profile()
with torch.no_grad():
    pr.enable()
    for _ in range(50): model(*args)
    pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
