#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_trainer.py tests/test_gpu_reference_goldens.py -q -m gpu -s -k "not arch_cases and not legacy" 2>&1 | grep -vE "^\s*$" | tail -60 | tee gpurun_out/r04k_pytest_tol.log
