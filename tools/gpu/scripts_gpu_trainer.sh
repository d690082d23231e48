#!/bin/bash
# device-resident train step: parity tests, then bench (device vs host trainer)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer.py -x -q -s 2>&1 | tail -40 | tee gpurun_out/pytest_trainer.log
rm -f gpurun_out/train.jsonl gpurun_out/train.err
for b in 138 207; do timeout 200 python bench.py --mode train --batch $b --steps 100 >> gpurun_out/train.jsonl 2>>gpurun_out/train.err; done
timeout 200 python bench.py --mode train --batch 207 --trainer host --steps 20 >> gpurun_out/train.jsonl 2>>gpurun_out/train.err
timeout 200 python bench.py --mode train --batch 2048 --steps 20 >> gpurun_out/train.jsonl 2>>gpurun_out/train.err
python - <<'PY'
import json
for l in open('gpurun_out/train.jsonl'):
    j=json.loads(l); print(j['config'].get('trainer'), j['config']['per_gpu_batch'], '%.3f ms/step' % j['ms_per_step'], '%.0f patches/s' % j['value'], 'e2e', j['e2e'] and round(j['e2e']['value']), 'loss/dim', j['config'].get('loss_per_dim_last_step'))
PY
tail -5 gpurun_out/train.err
