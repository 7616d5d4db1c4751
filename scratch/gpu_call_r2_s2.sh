#!/bin/bash
# session-2 validation: full GPU suite, smoke, ncu of the new frontier / ring kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/s2_gpu_tests.log 2>&1; tail -3 gpurun_out/s2_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'frontier_compact_kernel|ring_query_tma_kernel' -c 12 -o gpurun_out/s2_secondary python bench_rows.py --rows stream,ring > gpurun_out/s2_ncu_rows.log 2>&1; tail -2 gpurun_out/s2_ncu_rows.log | cut -c1-300
