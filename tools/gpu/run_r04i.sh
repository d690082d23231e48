#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r04i_bench_default.json 2>gpurun_out/r04i_bench_default.err; tail -3 gpurun_out/r04i_bench_default.err
python - <<'PY'
import json
j = json.load(open('gpurun_out/r04i_bench_default.json'))
print('value %.3f M  e2e %.3f M  frac %.4f fp32 %.3f' % (j['value']/1e6, j['e2e']['value']/1e6, j['roofline']['frac'], j['roofline_fp32']['frac']))
for k, v in (j.get('also') or {}).items():
    r = v['roofline']
    print('  %-32s %10.3f k patches/s  %.3f ms  roofline %s frac %.3f' % (k, v['value']/1e3, v['ms_per_step'], r['bound'], r['frac'] or -1), (' fp32 %.3f' % v['roofline_fp32']['frac']) if v.get('roofline_fp32') else '')
PY
