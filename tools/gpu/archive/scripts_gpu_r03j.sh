#!/bin/bash
# round-3j: cooperative small-batch batch-statistics chain (wrapper default mode)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_goldens.py -q -x 2>&1 | grep -n "^E \|passed\|failed\|FAILED\|Error" | head -20 | tee gpurun_out/pytest_r03j.log
timeout 300 python tools/gpu/wrapper_latency.py 2>&1 | tail -16 | tee gpurun_out/r03_wrapper_latency.txt
