#!/bin/bash
# parity of the fused chain + quick throughput numbers (no ncu)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_metrics.py -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_parity.log
rm -f gpurun_out/perf.jsonl gpurun_out/perf.err
timeout 120 python bench.py --no-cpu-baseline --no-e2e --steps 30 >> gpurun_out/perf.jsonl 2>>gpurun_out/perf.err
timeout 120 python bench.py --mode sample --no-cpu-baseline --no-e2e --steps 30 >> gpurun_out/perf.jsonl 2>>gpurun_out/perf.err
timeout 120 python bench.py --batch 1048576 --no-cpu-baseline --no-e2e --steps 3 >> gpurun_out/perf.jsonl 2>>gpurun_out/perf.err
python - <<'PY'
import json
for l in open('gpurun_out/perf.jsonl'):
    j=json.loads(l); print(j['metric'], j['config']['per_gpu_batch'], '%.3fM/s' % (j['value']/1e6), 'kernel_ms %.3f' % j['roofline']['kernel_ms'], 'fp32frac %.3f' % j['roofline_fp32']['frac'], j['clocks']['sm_mhz'])
PY
tail -q -n 3 gpurun_out/perf.err
