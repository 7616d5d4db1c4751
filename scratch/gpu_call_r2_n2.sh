#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
python - <<'PY'
import json
lines = [l for l in open('gpurun_out/n2_bench.json') if l.startswith('{')]
print(len(lines), 'json lines')
d = json.loads(lines[-1])
print('value', d['value'], 'n', d['n_gpus'], 'e2e', d['e2e']['value'], 'collective', d.get('collective'))
PY
grep "NCCL INFO" gpurun_out/n2_bench.err | grep -i "nranks\|Init COMPLETE" | head -3 | cut -c1-200; true
