#!/bin/bash
# 4 GPUs: bench.py under torchrun (collective block, NCCL log on stderr), then the loader micro-profile on one GPU
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/d_bench_n4.json 2> gpurun_out/d_bench_n4.err
tail -c 4000 gpurun_out/d_bench_n4.json; grep -i "nranks\|NVLS\|error\|Traceback" gpurun_out/d_bench_n4.err | head -12; tail -5 gpurun_out/d_bench_n4.err
python scratch/prof_loader.py 2>&1 | grep "us/batch"
