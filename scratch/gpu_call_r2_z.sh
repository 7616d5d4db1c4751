#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench_configs.py --config 3 2>/dev/null | cut -c1-200
timeout 200 python bench_configs.py --config 3 --lazy-edge-x 2>/dev/null | cut -c1-200
timeout 400 python -m pytest tests/test_gpu_attn_folded.py tests/test_gpu_lazy_edge_x.py -q > gpurun_out/z_tests.log 2>&1; tail -3 gpurun_out/z_tests.log
