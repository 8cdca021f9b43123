#!/bin/bash
# Four-GPU visit (charged 4x): weak scaling with the gather inside `value`, per-rank times, and GPU 0 alone.
TAG=${1:-r2n4}
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4"
FLAGS="--skip-cpu-baseline --skip-decoder --skip-train --skip-torch-gpu-baseline"
show() {
python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus')}); print((d.get('gather') or {}).get('by_rank')); print(d['e2e']['value'], d['e2e']['ms_per_step'])" 2>&1 | cut -c1-1500
}
echo "=== N=4 weak, fp32 gather" | tee $O/round_$TAG.log
timeout 600 $RUN --steps 5 --warmup 3 $FLAGS > $O/bench_${TAG}_weak.json 2> $O/bench_${TAG}_weak.err
show $O/bench_${TAG}_weak.json | tee -a $O/round_$TAG.log
tail -2 $O/bench_${TAG}_weak.err | tee -a $O/round_$TAG.log
echo "=== N=1 on the same box" | tee -a $O/round_$TAG.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 $FLAGS > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}_n1.err
show $O/bench_${TAG}_n1.json | tee -a $O/round_$TAG.log
