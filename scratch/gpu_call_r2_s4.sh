#!/bin/bash
# compute-sanitizer over the kernels added in session 2 (frontier segments, ring bulk-copy query, flat Time2Vec)
mkdir -p gpurun_out
SEL='test_frontier_compact[ or two_streams or large_ring or time2vec or bulk_ring'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/s4_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/s4_$tool.log | tail -3
done
