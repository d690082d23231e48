#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -s -k "not arch_cases and not legacy" 2>&1 | grep -vE "^\s*$" | tail -12 | tee gpurun_out/r04j_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --no-also --steps 30 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('value %.3f M  e2e %.3f M' % (j['value']/1e6, j['e2e']['value']/1e6))"
timeout 300 python bench.py --no-cpu-baseline --no-also --steps 30 --mode sample 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('sample value %.3f M  e2e %.3f M' % (j['value']/1e6, j['e2e']['value']/1e6))"
