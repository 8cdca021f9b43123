#!/bin/bash
# Round-2 GPU visit: the whole GPU suite (cascade wrapper test included), parity report, bench with the training leg.
TAG=${1:-r2d}
O=gpurun_out
mkdir -p $O
echo "=== tests" | tee $O/round_$TAG.log
timeout 2400 python -m pytest -q -m gpu tests/ -s 2>&1 | grep -vE "^\s*$" | tail -60 | cut -c1-500 | tee -a $O/round_$TAG.log
echo "=== parity" | tee -a $O/round_$TAG.log
timeout 900 python tests/parity_report.py $O/parity_$TAG.json 2>&1 | tail -4 | cut -c1-2500 | tee -a $O/round_$TAG.log
echo "=== bench" | tee -a $O/round_$TAG.log
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
echo "bench exit $?" | tee -a $O/round_$TAG.log
python -c "
import json,sys
d=json.loads(open('$O/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity','adapter_training','torch_gpu_baseline','cpu_baseline')})
print(d['e2e']); print(d['kernels'])" 2>&1 | cut -c1-3000 | tee -a $O/round_$TAG.log
tail -3 $O/bench_$TAG.err | tee -a $O/round_$TAG.log
