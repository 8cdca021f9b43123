#!/bin/bash
# Runs the GPU parity tests group by group, each under its own process + timeout, so that one
# failing kernel does not hide the other results.  Stops at the first CUDA FAULT (a numerics
# assertion is not a fault): faults cost the shared pool a device reset.
# Usage (on the GPU box): bash tools/gpu_selftest.sh [outfile]
OUT=${1:-gpurun_out/selftest.log}
mkdir -p "$(dirname "$OUT")"
: > "$OUT"
fault() { grep -qiE "illegal memory|launch failure|unspecified launch|device-side assert|misaligned address|Xid|mbarrier timeout|CUDA error" "$1"; }
run() {
  echo "=== $*" | tee -a "$OUT"
  timeout 900 "$@" > /tmp/st_last.log 2>&1
  echo "exit $?" >> /tmp/st_last.log
  tail -30 /tmp/st_last.log | tee -a "$OUT"
  if fault /tmp/st_last.log; then echo "!!! CUDA fault detected, stopping" | tee -a "$OUT"; exit 3; fi
}
PT="python -m pytest -q -x -m gpu"
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv | tee -a "$OUT"
run $PT tests/test_gpu_ops.py -k "layernorm or im2col or adapter"
run $PT tests/test_gpu_ops.py -k "attention"
run python tools/gemm_probe.py
run $PT tests/test_gpu_ops.py -k "gemm_plain"
run $PT tests/test_gpu_ops.py -k "gemm and not gemm_plain"
run $PT tests/test_gpu_encoder.py -k "golden and not vit_h"
run $PT tests/test_gpu_encoder.py -k "per_block or global_only or round_trip or refuses"
run $PT tests/test_gpu_encoder.py -k "vit_h"
echo "=== selftest complete" | tee -a "$OUT"
