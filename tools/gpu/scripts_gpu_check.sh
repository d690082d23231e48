#!/bin/bash
# whole GPU suite + the default bench line (what the driver runs) + sample mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 python bench.py --mode sample --no-cpu-baseline > gpurun_out/bench_sample.json 2> gpurun_out/bench_sample.err
python - <<'PY'
import json
for f in ('bench_default','bench_sample'):
    try:
        j=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, '%.3fM/s' % (j['value']/1e6), 'kernel_ms %.3f' % j['roofline']['kernel_ms'], 'fp32frac %.3f' % j['roofline_fp32']['frac'], 'e2e', j['e2e'] and '%.3fM/s' % (j['e2e']['value']/1e6), j.get('cpu_baseline'), j['clocks'])
    except Exception as e: print(f, 'FAILED', e)
PY
tail -q -n 3 gpurun_out/bench_default.err gpurun_out/bench_sample.err
