#!/bin/bash
timeout 300 python scratch/loader_probe.py 2>&1 | grep -v "^$" | cut -c1-160 | head -70
