#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_linear.py -q > gpurun_out/t_tc.log 2>&1; tail -15 gpurun_out/t_tc.log
timeout 120 python scratch/tc_time.py
