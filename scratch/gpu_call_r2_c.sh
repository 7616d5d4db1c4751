#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/c_gpu_tests.log 2>&1; tail -15 gpurun_out/c_gpu_tests.log
python bench.py > gpurun_out/c_bench_n1.json 2> gpurun_out/c_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','build_s')}, d['roofline']['frac'], d['e2e']['value'], d['loader_api']['us_per_batch'], d['loader_api']['with_negatives']['us_per_batch'], d['full_pass'])
PY
tail -3 gpurun_out/c_bench_n1.err
# launch list of the bench (shares of the step), then one full capture of the sampler in the steady state
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c_launches_bench_steps5.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:csr_sample_tma_kernel -s 3 -c 1 -o gpurun_out/c_sampler -f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c_ncu_full.log 2>&1; tail -2 gpurun_out/c_ncu_full.log
ls -la gpurun_out/*.ncu-rep
