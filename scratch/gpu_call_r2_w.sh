#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/w_gpu_tests.log 2>&1; tail -5 gpurun_out/w_gpu_tests.log
python bench_rows.py --rows dygformer 2>&1 | cut -c1-180
python bench_configs.py --config 5 2>/dev/null | cut -c1-400
