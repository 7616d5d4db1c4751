#!/bin/bash
# end-of-round refresh: bench (N=1), reference arm, rows, configs 3-5 on one fresh box
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/f_bench.json'))
print('build_s', d['build_s'], 'incl_build', d['full_pass']['value_incl_build'], 'value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'loader', d['loader_api']['us_per_batch'], d['loader_api']['with_negatives']['us_per_batch'], 'clocks', d['clocks'])
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/f_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/f_bench_ref.json
timeout 900 python bench_rows.py > gpurun_out/f_rows.jsonl 2> gpurun_out/f_rows.err; cut -c1-220 gpurun_out/f_rows.jsonl
for c in 3 4 5; do timeout 600 python bench_configs.py --config $c > gpurun_out/f_config$c.json 2> gpurun_out/f_config$c.err; cut -c1-300 gpurun_out/f_config$c.json; done
