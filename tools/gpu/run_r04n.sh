#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py -q -m gpu -s -k "gradients or adam or refused" 2>&1 | grep -vE "^\s*$" | tail -40 | tee gpurun_out/r04n_pytest_wide_train.log
