#!/bin/bash
# training path with the forward's log-sum-exp kept for the backward kernels: tests + training-step bench
tag=${1:-r2w}
out=gpurun_out/round_$tag.log
mkdir -p gpurun_out
{
echo "=== backward tests"
timeout 1200 python -m pytest tests/test_gpu_backward.py -x -q -m gpu 2>&1 | tail -6
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "=== training step"
timeout 900 python tools/train_bench.py --batch 8 2>&1 | tail -1 | tee gpurun_out/train_bench_$tag.json | cut -c1-1500
} > $out 2>&1
cat $out
