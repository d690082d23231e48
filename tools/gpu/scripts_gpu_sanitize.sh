#!/bin/bash
# compute-sanitizer passes over every kernel family (small batch); logs under gpurun_out/
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in ${NF_SANITIZE_TOOLS:-memcheck racecheck initcheck}; do
  NF_SANITIZE_N=37 timeout 900 $CS --tool $tool --print-limit 20 python tools/gpu/sanitize_smoke.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|^wide|^device|Error|hazard" gpurun_out/sanitize_$tool.log | head -8
done
