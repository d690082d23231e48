#!/bin/bash
# round-3n: chain-rule / Adam kernels parallel over (op, row) work items
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_train.py tests/test_gpu_reference_goldens.py -q -k "not arch_cases" 2>&1 | tail -2
rm -f gpurun_out/r03n.jsonl
for cfg in "138 0 1" "207 0 1" "296 0 1" "138 0 0" "2048 0 1"; do set -- $cfg; timeout 200 python bench.py --mode train --batch $1 --cta-warps $2 --fused $3 --steps 300 --no-cpu-baseline >> gpurun_out/r03n.jsonl 2>/dev/null; done
python - <<'PY'
import json
for l in open('gpurun_out/r03n.jsonl'):
    j=json.loads(l); print('batch', j['config']['per_gpu_batch'], 'fused', j['config'].get('fused'), '%.3f ms/step' % j['ms_per_step'], '%.0f patches/s' % j['value'], 'e2e', j['e2e'] and round(j['e2e']['value']))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03_launches_train_138_fused.csv python bench.py --mode train --batch 138 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_train.log 2>&1
grep -E "td_chain|td_apply|td_step|nf_reduce" gpurun_out/r03_launches_train_138_fused.csv | tail -4 | cut -d, -f5,12- | cut -c1-200
