#!/bin/bash
# Attention iteration visit: attention op tests, then op-level timings for the current kernels (per exp2
# polynomial share) and the mma.sync fallback kernel.
TAG=${1:-a}
O=gpurun_out
mkdir -p $O
L=$O/attn_$TAG.log
fault() { grep -qiE "illegal memory|launch failure|unspecified launch|misaligned address|mbarrier timeout|CUDA error" "$1"; }
echo "=== attention op tests" | tee $L
timeout 900 python -m pytest -q -m gpu tests/test_gpu_ops.py -k attention 2>&1 | tail -25 | tee -a $L
if fault $L; then echo "fault, stopping" | tee -a $L; exit 3; fi
for P in ${3:-0 2 4 6 8}; do
  echo "=== op bench (poly $P)" | tee -a $L
  CVLM_ATTN_POLY=$P timeout 600 python tools/op_bench.py ${2:-8} attn 2>&1 | tail -3 | tee -a $L
  if fault $L; then echo "fault, stopping" | tee -a $L; exit 3; fi
done
echo "=== op bench (mma.sync kernel)" | tee -a $L
CVLM_ATTN_LEGACY=1 timeout 600 python tools/op_bench.py ${2:-8} attn 2>&1 | tail -3 | tee -a $L
