#!/bin/bash
# ncu --set full capture of the qkv / proj / lin1 / lin2 GEMM launches of one block inside a B=32 bench step.
TAG=${1:-g}
O=gpurun_out
mkdir -p $O
BENCH="python bench.py --steps 1 --warmup 3 --batch 32 --skip-cpu-baseline --skip-e2e --skip-decoder"
$BENCH > $O/plain_gemm_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain_gemm_$TAG.log; exit 1; }
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_kernel -s 448 -c 4 \
    -o $O/prof_gemm_$TAG -f $BENCH > $O/ncu_gemm_$TAG.log 2>&1
echo "ncu gemm exit $?"; tail -2 $O/ncu_gemm_$TAG.log
