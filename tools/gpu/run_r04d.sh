#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu/wide_tc_check.py 256 512 2>&1 | tee gpurun_out/r04d_wide_tcs_check.log
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_reference_goldens.py -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/r04d_pytest_wide.log
