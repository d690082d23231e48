#!/bin/bash
# timing experiments on the hybrid kernel (library built with -DNFH_EXPERIMENTS): NF_HYB_EXP bit 0 = no accumulator wait,
# bit 1 = no fence.proxy.async; results of EXP != 0 may be wrong, only the time matters
for e in ${@:-0 1 2 3}; do
  NF_HYB_EXP=$e timeout 150 python bench.py --mode log_prob --tc 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-also 2>/dev/null | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('EXP $e  %.3f M patches/s  %.3f ms  nll/dim %s' % (j['value'] / 1e6, j['ms_per_step'], j.get('mean_nll_per_dim')))"
done
