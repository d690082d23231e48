#!/bin/bash
# quick check of the default chain kernel: bash tools/gpu/ab_quick.sh  -> log_prob and sample at 65 536 patches
for mode in log_prob sample; do
  timeout 150 python bench.py --mode $mode --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-also 2>/dev/null | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('$mode %.3f M patches/s  %.3f ms  nll/dim %s' % (j['value'] / 1e6, j['ms_per_step'], j.get('mean_nll_per_dim')))"
done
