#!/bin/bash
# round-3a: transposed warp reductions in the trainer passes + 4-slot host pipeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_train.py tests/test_gpu_reference_goldens.py -q -k "not arch_cases" 2>&1 | tail -25 | tee gpurun_out/pytest_r03a.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "host_buffer or full_size" 2>&1 | tail -3 | tee -a gpurun_out/pytest_r03a.log
rm -f gpurun_out/r03a.jsonl gpurun_out/r03a.err
for b in 138 207 2048; do timeout 200 python bench.py --mode train --batch $b --steps 100 --no-cpu-baseline >> gpurun_out/r03a.jsonl 2>>gpurun_out/r03a.err; done
timeout 300 python bench.py --mode sample --no-cpu-baseline >> gpurun_out/r03a.jsonl 2>>gpurun_out/r03a.err
timeout 300 python bench.py --no-cpu-baseline >> gpurun_out/r03a.jsonl 2>>gpurun_out/r03a.err
python - <<'PY'
import json
for l in open('gpurun_out/r03a.jsonl'):
    j=json.loads(l); print(j['metric'], 'batch', j['config']['per_gpu_batch'], '%.3f ms/step' % j['ms_per_step'], '%.0f patches/s' % j['value'], 'e2e', j['e2e'] and round(j['e2e']['value']))
PY
tail -q -n 3 gpurun_out/r03a.err
