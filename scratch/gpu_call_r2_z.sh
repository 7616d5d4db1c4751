#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/full_bench.json'))
print('build_s', d['build_s'], 'incl_build', d['full_pass']['value_incl_build'], 'value', d['value'], 'e2e', d['e2e']['value'])
PY
