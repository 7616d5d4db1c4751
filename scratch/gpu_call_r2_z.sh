#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_nn.py -q -k dygformer 2>&1 | tail -3
timeout 300 python scratch/dyg_ab.py 2>&1 | cut -c1-200
