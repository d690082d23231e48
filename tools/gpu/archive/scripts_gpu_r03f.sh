#!/bin/bash
# round-3f: whole GPU suite + compute-sanitizer (incl. synccheck) over all kernel families with the fused trainer
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_full.log
NF_SANITIZE_TOOLS="memcheck racecheck synccheck initcheck" bash tools/gpu/scripts_gpu_sanitize.sh 2>&1 | tee gpurun_out/sanitize_summary.log
