#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_reference_goldens.py -q -m gpu -k "wide or width" 2>&1 | tail -5
for w in 32 64 128 256 512; do b=16384; [ $w -ge 256 ] && b=4096; timeout 300 python bench.py --width $w --batch $b --steps 10 --no-cpu-baseline --no-e2e 2>>gpurun_out/r04m.err; done | tee gpurun_out/r04m_bench_wide.jsonl | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print('width', j['config']['width'], '%.3f ms' % j['ms_per_step'], '%.3f M patches/s' % (j['value'] / 1e6), 'tensor issued frac %.3f' % j['roofline']['issued_frac'])
"
tail -3 gpurun_out/r04m.err
