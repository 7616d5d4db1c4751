#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/full_gpu_tests.log 2>&1; tail -6 gpurun_out/full_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; tail -c 1500 gpurun_out/full_bench.json
