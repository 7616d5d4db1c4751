#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_linear.py -q > gpurun_out/f_tc.log 2>&1; tail -40 gpurun_out/f_tc.log
timeout 600 python -m pytest tests/test_gpu_nn.py -q -x > gpurun_out/f_nn.log 2>&1; tail -5 gpurun_out/f_nn.log
timeout 300 python bench_rows.py --rows dygformer 2>&1 | cut -c1-700
