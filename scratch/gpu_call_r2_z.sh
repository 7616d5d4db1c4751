#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_linear.py tests/test_gpu_attn_folded.py tests/test_gpu_nn.py tests/test_gpu_lazy_edge_x.py -q 2>&1 | tail -2
timeout 200 python bench_rows.py --rows tgat 2>&1 | cut -c1-230
timeout 200 python bench_configs.py --config 3 2>/dev/null | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv --log-file gpurun_out/x_tgat_launches.csv python scratch/tgat_probe.py 3 > /dev/null 2>&1
python scratch/launch_list.py gpurun_out/x_tgat_launches.csv 3 | head -3
