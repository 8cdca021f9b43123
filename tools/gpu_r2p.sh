#!/bin/bash
# Round-2 probe visit: bottleneck probes of the two attention kernels + validation of the plain-attention tile skip.
tag=${1:-r2p}
out=gpurun_out/round_$tag.log
mkdir -p gpurun_out
{
echo "=== plain attention tile skip"
timeout 600 python -m pytest tests/test_gpu_alpha_clip.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/tower_bench.py --batch 32 2>&1 | tail -1 | tee gpurun_out/tower_bench_$tag.json
echo "=== attention probes (batch 32)"
for f in g8 g9 g11; do
  timeout 120 python tools/attn_probe.py probes/libcvlm_$f.so 32 2>&1 | tail -1
done
} > $out 2>&1
cat $out
