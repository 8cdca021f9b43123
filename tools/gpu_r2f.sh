#!/bin/bash
TAG=${1:-r2f}
O=gpurun_out
mkdir -p $O
echo "=== backward tests" | tee $O/round_$TAG.log
timeout 1200 python -m pytest -q -m gpu tests/test_gpu_backward.py -x 2>&1 | grep -vE "^\s*$" | tail -25 | cut -c1-400 | tee -a $O/round_$TAG.log
echo "=== train step breakdown" | tee -a $O/round_$TAG.log
timeout 600 python tools/train_bench.py --batch 8 > $O/train_bench_$TAG.json 2> $O/train_bench_$TAG.err
cat $O/train_bench_$TAG.json | cut -c1-2500 | tee -a $O/round_$TAG.log
tail -3 $O/train_bench_$TAG.err | tee -a $O/round_$TAG.log
