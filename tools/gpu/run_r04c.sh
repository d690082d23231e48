#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_reference_goldens.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r04c_pytest_wide.log
for w in 32 64 128; do timeout 300 python bench.py --width $w --batch 16384 --steps 10 --no-cpu-baseline --no-e2e 2>>gpurun_out/r04c_wide.err; done | tee gpurun_out/r04c_bench_wide.jsonl | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print('width', j['config']['width'], '%.3f ms' % j['ms_per_step'], '%.3f M patches/s' % (j['value'] / 1e6))
"
tail -3 gpurun_out/r04c_wide.err
