#!/bin/bash
# op tests + op bench + encoder tests + bench (+ optional sweep)
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
bash tools/gpu_quick.sh $TAG bench
if [ "$2" == "sweep" ]; then
  echo "=== sweep" | tee -a $O/quick_$TAG.log
  timeout 1200 python tools/sweep.py $O/sweep_$TAG.jsonl 64 2>&1 | tail -50 | tee -a $O/quick_$TAG.log
fi
