#!/bin/bash
# round-3b: 8 vs 16 warps per patch-CTA after the cheaper reductions; ncu launch list of the train step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer.py -q 2>&1 | tail -8 | tee gpurun_out/pytest_r03b.log
rm -f gpurun_out/r03b.jsonl gpurun_out/r03b.err
for cfg in "138 8" "138 16" "138 0" "207 8" "207 16" "296 8" "64 8" "64 16"; do set -- $cfg; timeout 200 python bench.py --mode train --batch $1 --cta-warps $2 --steps 200 --no-cpu-baseline --no-e2e >> gpurun_out/r03b.jsonl 2>>gpurun_out/r03b.err; done
python - <<'PY'
import json
for l in open('gpurun_out/r03b.jsonl'):
    j=json.loads(l); print('batch', j['config']['per_gpu_batch'], 'cta_warps', j['config'].get('cta_warps'), '%.3f ms/step' % j['ms_per_step'], '%.0f patches/s' % j['value'])
PY
tail -q -n 3 gpurun_out/r03b.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03_launches_train_138.csv python bench.py --mode train --batch 138 --cta-warps 8 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_train_138.log 2>&1
tail -2 gpurun_out/ncu_train_138.log | cut -c1-300
