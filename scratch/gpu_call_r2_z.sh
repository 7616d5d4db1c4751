#!/bin/bash
timeout 200 python scratch/tgat_cpu_probe.py 2>&1 | cut -c1-150 | head -60
