#!/bin/bash
# round 2, call B: GPU suite with the new forms, build trace, bench, loader profile, ncu of the sampler
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/b_gpu_tests.log 2>&1; tail -40 gpurun_out/b_gpu_tests.log
TGM_B200_TRACE=1 python bench.py > gpurun_out/b_bench_n1.json 2> gpurun_out/b_bench_n1.err; tail -c 5000 gpurun_out/b_bench_n1.json; tail -20 gpurun_out/b_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/b_bench_ref.json 2> gpurun_out/b_bench_ref.err; tail -c 1500 gpurun_out/b_bench_ref.json; tail -3 gpurun_out/b_bench_ref.err
python scratch/prof_loader.py > gpurun_out/b_prof_loader.log 2>&1; grep "us/batch" gpurun_out/b_prof_loader.log
