#!/bin/bash
# First GPU call of the next round (one box, ~6 min):  gpurun --timeout 900 -- 'bash scratch/next_gpu_call.sh'
# Everything lands in gpurun_out/; copy what should be judged into profiles/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
# first hardware run of the MeanAggregator kernels (own process; remove the env gate once green)
TGM_B200_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_zz_gpu_tgn_mean.py -q > gpurun_out/gpu_tgn_mean.log 2>&1; tail -15 gpurun_out/gpu_tgn_mean.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
# config 4 (TGN) on one GPU, forward and training step: not measured in round 1
python bench_configs.py --config 4 > gpurun_out/config4_tgn_n1.json 2>> gpurun_out/configs.err
python bench_configs.py --config 4 --train > gpurun_out/config4_tgn_train_n1.json 2>> gpurun_out/configs.err
cat gpurun_out/config4_tgn_n1.json gpurun_out/config4_tgn_train_n1.json
# launch list of the training step (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_config4_train.csv \
    python bench_configs.py --config 4 --train --batches 5 > /dev/null 2>&1
# config 5 training step: eager launches vs the whole step replayed as one CUDA graph
python bench_configs.py --config 5 --train --batches 100 --edges 20000000 > gpurun_out/config5_train_eager.json 2>> gpurun_out/configs.err
timeout 300 python bench_configs.py --config 5 --train --cuda-graph --batches 100 --edges 20000000 > gpurun_out/config5_train_graph.json 2>> gpurun_out/configs.err
cat gpurun_out/config5_train_eager.json gpurun_out/config5_train_graph.json; tail -5 gpurun_out/configs.err
# config 3 training step: eager vs CUDA graph of the model step
python bench_configs.py --config 3 --train --batches 100 > gpurun_out/config3_train_eager.json 2>> gpurun_out/configs.err
timeout 300 python bench_configs.py --config 3 --train --cuda-graph --batches 100 > gpurun_out/config3_train_graph.json 2>> gpurun_out/configs.err
cat gpurun_out/config3_train_eager.json gpurun_out/config3_train_graph.json; tail -5 gpurun_out/configs.err
# (separate call, gpurun --gpus 4) config 4 time-sharded TGN, inference and training:
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench_tgn_shard.py --batches 500
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench_tgn_shard.py --batches 300 --train
# first hardware run of the reference-exact uniform sampling kernels (own process)
TGM_B200_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_zz_gpu_uniform_exact.py -q > gpurun_out/gpu_uniform_exact.log 2>&1; tail -8 gpurun_out/gpu_uniform_exact.log
