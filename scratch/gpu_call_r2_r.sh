#!/bin/bash
# 8 GPUs: bench.py exactly as the driver launches it
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r_bench_n8.json 2> gpurun_out/r_bench_n8.err
wc -l gpurun_out/r_bench_n8.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r_bench_n8.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value %.3e'%d['value'],'frac',round(d['roofline']['frac'],3),'e2e %.3e'%d['e2e']['value'],'link',d['e2e']['link'])
print('variants',{k:'%.3e'%v['value'] for k,v in d['e2e']['variants'].items()})
print('collective',d['collective'])
print('loader',d['loader_api']['us_per_batch'], d['loader_api']['with_negatives']['us_per_batch'])
PY
grep -c "NCCL INFO" gpurun_out/r_bench_n8.err; grep -i "nranks" gpurun_out/r_bench_n8.err | head -3; grep -i "error\|Traceback" gpurun_out/r_bench_n8.err | head -5
