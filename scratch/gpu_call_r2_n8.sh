#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err
python - <<'PY'
import json
lines = [l for l in open('gpurun_out/n8_bench.json') if l.startswith('{')]
print(len(lines), 'json lines')
d = json.loads(lines[-1])
print('value', d['value'], 'n', d['n_gpus'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['link'], 'fused', d['e2e']['variants']['fused_mean']['value'], 'collective', {k: d['collective'][k] for k in ('ms', 'algorithmic_gbs_per_gpu', 'dense_allreduce_ms', 'equals_dense_form')})
PY
