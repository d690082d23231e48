#!/bin/bash
# round-3d: tests of the fused step + ncu source-level capture of td_step_kernel at batch 138
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_train.py -q 2>&1 | tail -30 | tee gpurun_out/pytest_r03d.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:td_step_kernel -c 1 -f -o gpurun_out/prof_td_step_138 python bench.py --mode train --batch 138 --cta-warps 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_td_step.log 2>&1
tail -3 gpurun_out/ncu_td_step.log | cut -c1-300
ls -la gpurun_out/prof_td_step_138.ncu-rep
