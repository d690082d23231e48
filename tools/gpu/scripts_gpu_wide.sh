#!/bin/bash
# wide coupling nets: parity tests + throughput per width, then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_trainer.py -x -q -s 2>&1 | tail -40 | tee gpurun_out/pytest_wide.log
rm -f gpurun_out/wide.jsonl gpurun_out/wide.err
for w in 8 16 32; do timeout 300 python bench.py --width $w --batch 16384 --steps 10 --no-cpu-baseline --no-e2e >> gpurun_out/wide.jsonl 2>>gpurun_out/wide.err; done
python - <<'PY'
import json
for l in open('gpurun_out/wide.jsonl'):
    j=json.loads(l); print('width', j['config'].get('width'), '%.3f ms/step' % j['ms_per_step'], '%.3f M patches/s' % (j['value']/1e6), 'fp32 frac %.3f' % j['roofline_fp32']['frac'])
PY
tail -5 gpurun_out/wide.err
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
