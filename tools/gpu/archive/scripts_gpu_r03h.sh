#!/bin/bash
# round-3h: ncu source-level capture of the current fused step kernel (batch 138, 8 warps) + per-pass check at 2048
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:td_step_kernel -c 1 -f -o gpurun_out/prof_td_step_138_v2 python bench.py --mode train --batch 138 --cta-warps 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_td_step.log 2>&1
tail -2 gpurun_out/ncu_td_step.log | cut -c1-200
for cfg in "2048 0 1" "138 8 1" "138 16 1"; do set -- $cfg; timeout 200 python bench.py --mode train --batch $1 --cta-warps $2 --fused $3 --steps 200 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch', j['config']['per_gpu_batch'], 'cta_warps', j['config'].get('cta_warps'), '%.3f ms/step' % j['ms_per_step'])"; done
