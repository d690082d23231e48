#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/tc_probe tools/tc_probe.cu && timeout 60 gpurun_out/tc_probe 2>&1 | tee gpurun_out/tc_probe.log
timeout 300 python -m pytest tests/test_gpu_tensor_core.py -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_tc.log
timeout 120 python bench.py --tc 1 --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/bench_tc.json')); print('TC bench', '%.3fM/s' % (j['value']/1e6), 'kernel_ms %.3f' % j['roofline']['kernel_ms'], j['config']['mean_nll_per_dim'])
except Exception as e: print('bench_tc failed', e, open('gpurun_out/bench_tc.err').read()[-500:])
PY
