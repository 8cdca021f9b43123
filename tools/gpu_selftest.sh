#!/bin/bash
# Runs the GPU parity tests group by group, each under its own process + timeout, so that one
# faulting kernel (trap / illegal access kills the CUDA context) does not hide the other results.
# Usage (on the GPU box): bash tools/gpu_selftest.sh [outfile]
OUT=${1:-gpurun_out/selftest.log}
mkdir -p "$(dirname "$OUT")"
: > "$OUT"
run() {
  echo "=== $*" | tee -a "$OUT"
  timeout 600 python -m pytest -q -x -m gpu "$@" 2>&1 | tail -25 | tee -a "$OUT"
}
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv | tee -a "$OUT"
run tests/test_gpu_ops.py -k "layernorm or im2col or adapter"
run tests/test_gpu_ops.py -k "gemm_plain"
run tests/test_gpu_ops.py -k "gemm and not gemm_plain"
run tests/test_gpu_ops.py -k "attention"
run tests/test_gpu_encoder.py -k "golden and not vit_h"
run tests/test_gpu_encoder.py -k "per_block or global_only or round_trip or refuses"
run tests/test_gpu_encoder.py -k "vit_h"
