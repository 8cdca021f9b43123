#!/bin/bash
# Two-GPU visit: weak scaling (32 images per GPU) with the pipelined gather inside `value`, bf16 payload, and the
# strong-scaling line of BASELINE.json configs[2] (256 images in total).
TAG=${1:-r2n2}
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2"
echo "=== N=2 weak, fp32 gather" | tee $O/round_$TAG.log
timeout 900 $RUN --steps 5 --warmup 3 --skip-cpu-baseline --skip-decoder --skip-train > $O/bench_${TAG}_weak.json 2> $O/bench_${TAG}_weak.err
echo "exit $?" | tee -a $O/round_$TAG.log
python -c "
import json
d=json.loads(open('$O/bench_${TAG}_weak.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling','gather')}); print(d['e2e'])" 2>&1 | cut -c1-1500 | tee -a $O/round_$TAG.log
tail -3 $O/bench_${TAG}_weak.err | tee -a $O/round_$TAG.log
echo "=== N=2 weak, bf16 gather" | tee -a $O/round_$TAG.log
timeout 900 $RUN --steps 5 --warmup 3 --skip-cpu-baseline --skip-decoder --skip-train --gather-dtype bf16 > $O/bench_${TAG}_weak_bf16.json 2> $O/bench_${TAG}_weak_bf16.err
python -c "
import json
d=json.loads(open('$O/bench_${TAG}_weak_bf16.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gather')}); print(d['e2e'])" 2>&1 | cut -c1-1500 | tee -a $O/round_$TAG.log
echo "=== N=2 strong, 256 images in total" | tee -a $O/round_$TAG.log
timeout 900 $RUN --steps 3 --warmup 3 --skip-cpu-baseline --skip-decoder --skip-train --total-batch 256 > $O/bench_${TAG}_strong256.json 2> $O/bench_${TAG}_strong256.err
python -c "
import json
d=json.loads(open('$O/bench_${TAG}_strong256.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','scaling','gather')}); print(d['config']); print(d['e2e'])" 2>&1 | cut -c1-1800 | tee -a $O/round_$TAG.log
tail -3 $O/bench_${TAG}_strong256.err | tee -a $O/round_$TAG.log
echo "=== N=1 strong-scaling reference point (256 images on one GPU)" | tee -a $O/round_$TAG.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --skip-cpu-baseline --skip-decoder --skip-train --skip-torch-gpu-baseline --total-batch 256 > $O/bench_${TAG}_strong256_n1.json 2> $O/bench_${TAG}_strong256_n1.err
python -c "
import json
d=json.loads(open('$O/bench_${TAG}_strong256_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','scaling')}); print(d['e2e'])" 2>&1 | cut -c1-1200 | tee -a $O/round_$TAG.log
