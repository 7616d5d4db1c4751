#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/l_gpu_tests.log 2>&1; tail -6 gpurun_out/l_gpu_tests.log
python bench_rows.py --rows dygformer,tgat,twohop > gpurun_out/l_rows.jsonl 2>gpurun_out/l_rows.err; cut -c1-420 gpurun_out/l_rows.jsonl; tail -3 gpurun_out/l_rows.err
python bench_configs.py --config 3 > gpurun_out/l_config3.json 2>> gpurun_out/l_cfg.err; cat gpurun_out/l_config3.json | cut -c1-900
python bench_configs.py --config 5 > gpurun_out/l_config5.json 2>> gpurun_out/l_cfg.err; cat gpurun_out/l_config5.json | cut -c1-900
python bench_configs.py --config 4 > gpurun_out/l_config4.json 2>> gpurun_out/l_cfg.err; cat gpurun_out/l_config4.json | cut -c1-900; tail -5 gpurun_out/l_cfg.err
