#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_goldens.py tests/test_gpu_wide.py -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r04g_pytest.log
