#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:time2vec -s 2 -c 1 -o gpurun_out/t2v -f python scratch/t2v_probe.py > gpurun_out/t2v_ncu.log 2>&1; tail -2 gpurun_out/t2v_ncu.log
