#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_nn.py tests/test_gpu_tc_linear.py tests/test_gpu_attn_folded.py -q 2>&1 | tail -2
timeout 200 python bench_rows.py --rows dygformer,tgat 2>&1 | cut -c1-200
timeout 200 python bench_configs.py --config 5 2>/dev/null | cut -c1-260
