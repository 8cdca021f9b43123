#!/bin/bash
TAG=${1:-r2e}
O=gpurun_out
mkdir -p $O
echo "=== train step breakdown" | tee $O/round_$TAG.log
timeout 600 python tools/train_bench.py --batch 8 > $O/train_bench_$TAG.json 2> $O/train_bench_$TAG.err
cat $O/train_bench_$TAG.json | cut -c1-2500 | tee -a $O/round_$TAG.log
tail -3 $O/train_bench_$TAG.err | tee -a $O/round_$TAG.log
echo "=== sweep to batch 128" | tee -a $O/round_$TAG.log
timeout 1200 python tools/sweep.py $O/sweep_${TAG}_config5.jsonl 128 2>&1 | tail -50 | cut -c1-200 | tee -a $O/round_$TAG.log
