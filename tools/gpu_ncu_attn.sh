#!/bin/bash
# ncu --set full capture of one launch of each attention kernel from the op-level bench (B = $2, default 8).
# usage: bash tools/gpu_ncu_attn.sh <tag> [batch]
TAG=${1:-a}; B=${2:-8}
O=gpurun_out
mkdir -p $O
timeout 600 python tools/op_bench.py $B attn > $O/ncu_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $O/ncu_plain_$TAG.log; exit 1; }
cat $O/ncu_plain_$TAG.log
for K in attn_global_tc2_kernel attn_window_tc2_kernel; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o $O/prof_${K}_$TAG -f \
    python tools/op_bench.py $B attn > $O/ncu_${K}_$TAG.log 2>&1
echo "ncu $K exit $?"; tail -2 $O/ncu_${K}_$TAG.log
done
