#!/bin/bash
# host-buffer entry point with cross-call overlap: chunk sizes
tag=${1:-r2x}
out=gpurun_out/round_$tag.log
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_memory_hygiene.py -x -q -m gpu -k "host or hygiene or poison" 2>&1 | tail -3
for ch in 8 16 4; do
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --skip-decoder --skip-train --skip-torch-gpu-baseline --e2e-chunk $ch > gpurun_out/bench_${tag}_ch$ch.json 2> gpurun_out/bench_${tag}_ch$ch.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}_ch$ch.json').read().strip().splitlines()[-1])
print('chunk $ch', {k:d.get(k) for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['ms_per_step'])"
done
} > $out 2>&1
cat $out
