#!/bin/bash
# Quick GPU visit: op tests for the kernels under development, op-level timings, optional bench.
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
fault() { grep -qiE "illegal memory|launch failure|unspecified launch|misaligned address|mbarrier timeout|CUDA error" "$1"; }
echo "=== op tests" | tee $O/quick_$TAG.log
timeout 900 python -m pytest -q -x -m gpu tests/test_gpu_ops.py 2>&1 | tail -25 | tee -a $O/quick_$TAG.log
if fault $O/quick_$TAG.log; then echo "fault, stopping" | tee -a $O/quick_$TAG.log; exit 3; fi
echo "=== op bench" | tee -a $O/quick_$TAG.log
timeout 600 python tools/op_bench.py 8 2>&1 | tail -12 | tee -a $O/quick_$TAG.log
if fault $O/quick_$TAG.log; then echo "fault, stopping" | tee -a $O/quick_$TAG.log; exit 3; fi
if [ "$2" == "bench" ]; then
  echo "=== encoder tests" | tee -a $O/quick_$TAG.log
  timeout 900 python -m pytest -q -x -m gpu tests/test_gpu_encoder.py 2>&1 | tail -8 | tee -a $O/quick_$TAG.log
  echo "=== bench" | tee -a $O/quick_$TAG.log
  timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > $O/bench_$TAG.json 2> $O/bench_$TAG.err
  echo "bench exit $?" | tee -a $O/quick_$TAG.log
  python - <<PY | tee -a $O/quick_$TAG.log
import json
d=json.loads(open('$O/bench_$TAG.json').read().strip().splitlines()[-1])
tot=d['ms_per_step']; print('value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'ms/step',round(tot,2),'gemm TF/s',round(d['roofline']['achieved'],1))
for k,v in d['kernels'].items(): print(f"{k:14s} {v['ms_per_step']:8.2f} ms {v['launches_per_step']:4d} {100*v['ms_per_step']/tot:5.1f}%")
PY
fi
