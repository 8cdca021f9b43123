#!/bin/bash
# Eight-GPU visit (charged 8x): BASELINE.json configs[2] (256 images in total, batch-sharded 32 per GPU, embeddings
# gathered on every rank, gather inside `value`) with fp32 and bf16 payload, and the one-GPU point of the same box.
TAG=${1:-r2n8}
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8"
FLAGS="--skip-cpu-baseline --skip-decoder --skip-train --skip-torch-gpu-baseline"
show() {
python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling','gather','clocks')}); print(d['e2e'])" 2>&1 | cut -c1-1500
}
echo "=== N=8 weak, fp32 gather" | tee $O/round_$TAG.log
timeout 600 $RUN --steps 5 --warmup 3 $FLAGS > $O/bench_${TAG}_weak.json 2> $O/bench_${TAG}_weak.err
echo "exit $?" | tee -a $O/round_$TAG.log
show $O/bench_${TAG}_weak.json | tee -a $O/round_$TAG.log
tail -2 $O/bench_${TAG}_weak.err | tee -a $O/round_$TAG.log
echo "=== N=8 weak, bf16 gather" | tee -a $O/round_$TAG.log
timeout 600 $RUN --steps 5 --warmup 3 $FLAGS --gather-dtype bf16 > $O/bench_${TAG}_weak_bf16.json 2> $O/bench_${TAG}_weak_bf16.err
show $O/bench_${TAG}_weak_bf16.json | tee -a $O/round_$TAG.log
echo "=== N=1 on the same box" | tee -a $O/round_$TAG.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 $FLAGS > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}_n1.err
show $O/bench_${TAG}_n1.json | tee -a $O/round_$TAG.log
