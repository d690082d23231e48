#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_train.py -q 2>&1 | grep -v "^$" | grep -n "^E \|passed\|failed\|FAILED\|assert" | head -40 | tee gpurun_out/pytest_r03e.log
rm -f gpurun_out/r03e.jsonl gpurun_out/r03e.err
for cfg in "138 0 1" "138 8 1" "138 0 0" "207 0 1" "207 0 0" "296 0 1" "64 0 1"; do set -- $cfg; timeout 200 python bench.py --mode train --batch $1 --cta-warps $2 --fused $3 --steps 200 --no-cpu-baseline --no-e2e >> gpurun_out/r03e.jsonl 2>>gpurun_out/r03e.err; done
python - <<'PY'
import json
for l in open('gpurun_out/r03e.jsonl'):
    j=json.loads(l); print('batch', j['config']['per_gpu_batch'], 'cta_warps', j['config'].get('cta_warps'), 'fused', j['config'].get('fused'), '%.3f ms/step' % j['ms_per_step'], '%.0f patches/s' % j['value'], 'launches', j['gpu_launches'])
PY
tail -q -n 3 gpurun_out/r03e.err
