#!/bin/bash
# BASELINE config 4: NLL throughput vs batch size on one GPU (kernel-only value; no e2e / CPU arm)
mkdir -p gpurun_out
rm -f gpurun_out/sweep_batch.jsonl
for b in 1024 4096 16384 65536 262144 1048576; do
  steps=$(( 6553600 / b )); [ $steps -lt 3 ] && steps=3; [ $steps -gt 400 ] && steps=400
  timeout 300 python bench.py --batch $b --steps $steps --no-cpu-baseline --no-e2e >> gpurun_out/sweep_batch.jsonl 2>> gpurun_out/sweep_batch.err
done
python - <<'PY'
import json
print("| batch | M patches/s | ms/step | HBM frac | FP32 frac |")
print("|---|---|---|---|---|")
for l in open('gpurun_out/sweep_batch.jsonl'):
    j=json.loads(l); print("| %d | %.2f | %.3f | %.4f | %.3f |" % (j['config']['per_gpu_batch'], j['value']/1e6, j['ms_per_step'], j['roofline']['frac'], j['roofline_fp32']['frac']))
PY
