#!/bin/bash
# whole GPU suite without -x (every failure is listed) + smoke()
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/pytest_gpu_full.log
tail -15 gpurun_out/pytest_gpu_full.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
