#!/bin/bash
# stem / neck trims (adapter_embed, neck cast fused into the last lin2): full GPU suite + bench
tag=${1:-r2t}
out=gpurun_out/round_$tag.log
mkdir -p gpurun_out
{
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "=== bench"
timeout 900 python bench.py --steps 5 --warmup 3 --skip-decoder --skip-train > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','clocks','gpu_launches')}); print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}); print(d['parity']); print(d['e2e']['value']); print(d['roofline']['whole_step'])"
} > $out 2>&1
cat $out
