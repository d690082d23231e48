#!/bin/bash
# default chain kernel as one 16-warp CTA per SM vs two 8-warp CTAs (NF_WINO_CTA_WARPS=8), then the GPU suite in the 8-warp shape
mkdir -p gpurun_out
for w in 16 8; do echo "NF_WINO_CTA_WARPS=$w"; NF_WINO_CTA_WARPS=$w bash tools/gpu/ab_quick.sh; done 2>&1 | tee gpurun_out/ab_cta_warps.log
NF_WINO_CTA_WARPS=8 timeout 70 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee -a gpurun_out/ab_cta_warps.log
