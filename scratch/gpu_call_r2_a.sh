#!/bin/bash
# round 2, first call: whole GPU suite with the formerly gated tests, then baseline bench
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_tgn_mean.py --deselect tests/test_gpu_uniform_exact.py > gpurun_out/a_gpu_tests.log 2>&1; tail -3 gpurun_out/a_gpu_tests.log
timeout 300 python -m pytest tests/test_gpu_tgn_mean.py -q > gpurun_out/a_gpu_tgn_mean.log 2>&1; tail -30 gpurun_out/a_gpu_tgn_mean.log
timeout 300 python -m pytest tests/test_gpu_uniform_exact.py -q > gpurun_out/a_gpu_uniform_exact.log 2>&1; tail -30 gpurun_out/a_gpu_uniform_exact.log
python bench.py > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err; tail -c 1500 gpurun_out/a_bench_n1.json
