#!/bin/bash
mkdir -p gpurun_out
python scratch/frontier_probe.py 2>&1 | tail -12
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "frontier" 2>&1 | tail -2
for k in frontier_compact_kernel ring_query_tma_kernel time2vec_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/s3_$k python bench_rows.py --rows stream,ring > gpurun_out/s3_ncu_$k.log 2>&1; tail -1 gpurun_out/s3_ncu_$k.log | cut -c1-200
done
