#!/bin/bash
# window attention with part of the exponentials on the FMA pipe: correctness + W2_POLY sweep
tag=${1:-r2y}
out=gpurun_out/round_$tag.log
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" 2>&1 | tail -3
echo "shipped (W2_POLY=4):"
timeout 120 python tools/attn_probe.py camouflaged_vlm_b200/libcvlm_b200.so 32 2>&1 | tail -1
for f in $(ls probes/*.so); do
  timeout 120 python tools/attn_probe.py $f 32 2>&1 | tail -1
done
timeout 120 python tools/attn_probe.py camouflaged_vlm_b200/libcvlm_b200.so 8 2>&1 | tail -1
} > $out 2>&1
cat $out
