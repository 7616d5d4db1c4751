#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:attn_warp_kernel -s 3 -c 1 -o gpurun_out/y_warp -f python scratch/tgat_probe.py 2 > gpurun_out/y_ncu.log 2>&1
tail -3 gpurun_out/y_ncu.log
