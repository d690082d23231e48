#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -s 2>&1 | tail -30 | tee gpurun_out/pytest_train.log
