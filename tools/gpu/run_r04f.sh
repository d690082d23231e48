#!/bin/bash
# round-4f: exchange-buffer epilogue 3; N-batching variants of the tensor-core wide kernel; whole GPU suite; default bench line
mkdir -p gpurun_out
rm -f gpurun_out/r04f_variants.txt
for flags in "" "-DNF_WTC_NB3=0" "-DNF_WTC_NB2=0 -DNF_WTC_NB3=0"; do
  touch noise_flow_b200/csrc/nf_wide_tc.cu
  NF_EXTRA_NVCC_FLAGS="$flags" python -m noise_flow_b200.build > /dev/null 2>>gpurun_out/r04f.err
  for w in 32 64 128; do
    timeout 300 python bench.py --width $w --batch 16384 --steps 10 --no-cpu-baseline --no-e2e 2>>gpurun_out/r04f.err | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print('flags [$flags] width', j['config']['width'], '%.3f ms' % j['ms_per_step'], '%.3f M patches/s' % (j['value'] / 1e6), 'issued frac %.3f' % j['roofline'].get('issued_frac', -1))
" | tee -a gpurun_out/r04f_variants.txt
  done
done
touch noise_flow_b200/csrc/nf_wide_tc.cu; python -m noise_flow_b200.build > /dev/null
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r04f_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r04f_bench_default.json 2>gpurun_out/r04f_bench_default.err; tail -3 gpurun_out/r04f_bench_default.err
python - <<'PY'
import json
j = json.load(open('gpurun_out/r04f_bench_default.json'))
print('value %.3f M  e2e %.3f M  frac %.4f fp32 %.3f' % (j['value']/1e6, j['e2e']['value']/1e6, j['roofline']['frac'], j['roofline_fp32']['frac']))
for k, v in (j.get('also') or {}).items():
    r = v['roofline']
    print('  %-32s %10.3f k patches/s  %.3f ms  roofline %s frac %.3f' % (k, v['value']/1e3, v['ms_per_step'], r['bound'], r['frac'] or -1), (' fp32 %.3f' % v['roofline_fp32']['frac']) if v.get('roofline_fp32') else '')
PY
