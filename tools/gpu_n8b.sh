#!/bin/bash
# Eight-GPU diagnosis (charged 8x): per-rank encoder time vs time including the gather, default NCCL and NCCL limited to few CTAs.
TAG=${1:-r2n8b}
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8"
FLAGS="--skip-cpu-baseline --skip-decoder --skip-train --skip-torch-gpu-baseline"
show() {
python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step')}); print(d['gather'].get('by_rank')); print(d['e2e']['value'])" 2>&1 | cut -c1-1500
}
echo "=== N=8 weak, fp32 gather, default NCCL" | tee $O/round_$TAG.log
timeout 600 $RUN --steps 5 --warmup 3 $FLAGS > $O/bench_${TAG}_a.json 2> $O/bench_${TAG}_a.err
show $O/bench_${TAG}_a.json | tee -a $O/round_$TAG.log
echo "=== N=8 weak, fp32 gather, NCCL_MAX_CTAS=4" | tee -a $O/round_$TAG.log
NCCL_MAX_CTAS=4 timeout 600 $RUN --steps 5 --warmup 3 $FLAGS > $O/bench_${TAG}_b.json 2> $O/bench_${TAG}_b.err
show $O/bench_${TAG}_b.json | tee -a $O/round_$TAG.log
