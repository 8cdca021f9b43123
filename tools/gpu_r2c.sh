#!/bin/bash
TAG=${1:-r2c}
O=gpurun_out
mkdir -p $O
echo "=== grad diag" | tee $O/round_$TAG.log
timeout 600 python tools/grad_diag.py 2>&1 | tail -12 | tee -a $O/round_$TAG.log
echo "=== backward tests" | tee -a $O/round_$TAG.log
timeout 1500 python -m pytest -q -m gpu tests/test_gpu_backward.py -s -k "adapter or contract" 2>&1 | grep -vE "^\s*$" | tail -40 | cut -c1-600 | tee -a $O/round_$TAG.log
