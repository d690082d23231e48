#!/bin/bash
# quick GPU visit: parity tests + warps sweep (no ncu)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep.jsonl gpurun_out/sweep.err
for w in 12 14 16; do timeout 120 python bench.py --warps $w --no-cpu-baseline --no-e2e --steps 20 >> gpurun_out/sweep.jsonl 2>>gpurun_out/sweep.err; done
timeout 120 python bench.py --batch 4096 --no-cpu-baseline --no-e2e --steps 100 >> gpurun_out/sweep.jsonl 2>>gpurun_out/sweep.err
timeout 120 python bench.py --mode sample --no-cpu-baseline --no-e2e --steps 20 >> gpurun_out/sweep.jsonl 2>>gpurun_out/sweep.err
tail -3 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    j=json.loads(l); print(j['metric'], j['config']['per_gpu_batch'], '%.3fM/s' % (j['value']/1e6), 'kernel_ms %.3f' % j['roofline']['kernel_ms'], 'fp32frac %.3f' % j['roofline_fp32']['frac'], j['clocks']['sm_mhz'])
PY
tail -3 gpurun_out/sweep.err
