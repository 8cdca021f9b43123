#!/bin/bash
# Eight-GPU visit (charged 8x): the weak-scaling line with the gather inside `value`, and BASELINE.json configs[2] exactly
# (256 images in total, batch-sharded, embeddings gathered on every rank).
TAG=${1:-r2n8}
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8"
FLAGS="--skip-cpu-baseline --skip-decoder --skip-train"
echo "=== N=8 weak, fp32 gather" | tee $O/round_$TAG.log
timeout 600 $RUN --steps 5 --warmup 3 $FLAGS > $O/bench_${TAG}_weak.json 2> $O/bench_${TAG}_weak.err
echo "exit $?" | tee -a $O/round_$TAG.log
python -c "
import json
d=json.loads(open('$O/bench_${TAG}_weak.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling','gather','clocks')}); print(d['e2e'])" 2>&1 | cut -c1-1500 | tee -a $O/round_$TAG.log
tail -2 $O/bench_${TAG}_weak.err | tee -a $O/round_$TAG.log
echo "=== N=8 strong, 256 images in total (32 per GPU)" | tee -a $O/round_$TAG.log
timeout 600 $RUN --steps 5 --warmup 3 $FLAGS --total-batch 256 > $O/bench_${TAG}_strong256.json 2> $O/bench_${TAG}_strong256.err
python -c "
import json
d=json.loads(open('$O/bench_${TAG}_strong256.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','scaling','gather')}); print(d['e2e'])" 2>&1 | cut -c1-1500 | tee -a $O/round_$TAG.log
echo "=== N=8 weak, bf16 gather" | tee -a $O/round_$TAG.log
timeout 600 $RUN --steps 5 --warmup 3 $FLAGS --gather-dtype bf16 > $O/bench_${TAG}_weak_bf16.json 2> $O/bench_${TAG}_weak_bf16.err
python -c "
import json
d=json.loads(open('$O/bench_${TAG}_weak_bf16.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gather')}); print(d['e2e'])" 2>&1 | cut -c1-1200 | tee -a $O/round_$TAG.log
