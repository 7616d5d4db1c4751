#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/full_gpu_tests.log 2>&1; tail -3 gpurun_out/full_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/full_bench.json'))
print('build_s', d['build_s'], 'incl_build', d['full_pass']['value_incl_build'], 'value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'loader', d['loader_api']['us_per_batch'], d['loader_api']['with_negatives']['us_per_batch'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-400
timeout 600 python bench_rows.py > gpurun_out/full_rows.jsonl 2> gpurun_out/full_rows.err; cut -c1-150 gpurun_out/full_rows.jsonl
