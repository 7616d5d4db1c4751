#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/e_gpu_tests.log 2>&1; tail -4 gpurun_out/e_gpu_tests.log
python bench.py > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/e_bench_n1.json'))
print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['loader_api']['us_per_batch'], d['loader_api']['with_negatives']['us_per_batch'], d['eager_cuda_baseline'], d['cpu_baseline']['value'])
PY
tail -3 gpurun_out/e_bench_n1.err
python bench_rows.py --rows ring,tgat,tgn,dygformer,stream > gpurun_out/e_rows.jsonl 2> gpurun_out/e_rows.err; cat gpurun_out/e_rows.jsonl | cut -c1-1500; tail -3 gpurun_out/e_rows.err
