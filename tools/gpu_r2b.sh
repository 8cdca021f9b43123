#!/bin/bash
# Round-2 GPU visit: backward-pass tests, memory-hygiene tests, the changed encoder / parity tests.
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
echo "=== backward + hygiene tests" | tee $O/round_$TAG.log
timeout 1500 python -m pytest -q -m gpu tests/test_gpu_backward.py tests/test_gpu_memory_hygiene.py -s 2>&1 | grep -vE "^\s*$" | tail -80 | tee -a $O/round_$TAG.log
echo "=== encoder / parity tests" | tee -a $O/round_$TAG.log
timeout 1500 python -m pytest -q -m gpu tests/test_gpu_encoder.py tests/test_gpu_parity_b32.py tests/test_gpu_decoder.py tests/test_gpu_ops.py -s 2>&1 | grep -vE "^\s*$" | tail -40 | tee -a $O/round_$TAG.log
