#!/bin/bash
mkdir -p gpurun_out
timeout 100 python scratch/minb_ab.py 2>&1 | tail -4
timeout 400 python -m pytest tests/test_gpu_attn_folded.py -q > gpurun_out/z_tests.log 2>&1; tail -3 gpurun_out/z_tests.log
