#!/bin/bash
# compute-sanitizer over the kernels added this round
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_fused_forms.py tests/test_gpu_negatives.py tests/test_gpu_lazy_edge_x.py -k "not config1" > gpurun_out/o_memcheck_a.log 2>&1; echo "memcheck sampler forms rc=$?"; tail -4 gpurun_out/o_memcheck_a.log
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_tc_linear.py -k "tails or tiny or out_proj or alias" > gpurun_out/o_memcheck_b.log 2>&1; echo "memcheck tc_linear rc=$?"; tail -4 gpurun_out/o_memcheck_b.log
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_nn.py -k "dygformer" > gpurun_out/o_memcheck_c.log 2>&1; echo "memcheck dygformer rc=$?"; tail -4 gpurun_out/o_memcheck_c.log
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_fused_forms.py -k "fused_mean and k20" tests/test_gpu_nn.py::test_dygformer_matches_reference_module > gpurun_out/o_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/o_racecheck.log
