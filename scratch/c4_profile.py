import cProfile, pstats, sys, io
sys.path.insert(0, '.')
sys.argv = ['bench_configs.py', '--config', '4', '--batches', '700']
import bench_configs
pr = cProfile.Profile()
pr.enable()
try:
    bench_configs.main()
finally:
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22)
    print(s.getvalue()[:6000])
