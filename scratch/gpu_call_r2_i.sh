#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tc3_linear -s 1 -c 1 -o gpurun_out/i_tc -f python scratch/tc_probe.py > gpurun_out/i_ncu.log 2>&1; tail -3 gpurun_out/i_ncu.log
