#!/bin/bash
# global attention with two issuers and a separate P region: correctness + timing
tag=${1:-r2s}
out=gpurun_out/round_$tag.log
mkdir -p gpurun_out
{
echo "=== attention / tower / encoder tests"
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_alpha_clip.py tests/test_gpu_encoder.py tests/test_gpu_parity_b32.py -x -q -m gpu 2>&1 | tail -4
echo "=== timing"
timeout 120 python tools/attn_probe.py camouflaged_vlm_b200/libcvlm_b200.so 32 2>&1 | tail -1
timeout 120 python tools/attn_probe.py camouflaged_vlm_b200/libcvlm_b200.so 8 2>&1 | tail -1
timeout 300 python tools/tower_bench.py --batch 32 2>&1 | tail -1 | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --skip-decoder --skip-train --skip-torch-gpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','clocks')}); print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}); print(d['parity'])"
} > $out 2>&1
cat $out
