#!/bin/bash
# Two-GPU check of the host-buffer path with the per-chunk gather (Python choreography) after the cross-call overlap change.
TAG=${1:-r2n2b}
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2"
FLAGS="--skip-cpu-baseline --skip-decoder --skip-train --skip-torch-gpu-baseline"
echo "=== forward_host tests (one GPU)" | tee $O/round_$TAG.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_encoder.py -x -q -m gpu -k "host" 2>&1 | tail -2 | tee -a $O/round_$TAG.log
echo "=== N=2 weak, fp32 gather" | tee -a $O/round_$TAG.log
timeout 900 $RUN --steps 5 --warmup 3 $FLAGS > $O/bench_${TAG}_weak.json 2> $O/bench_${TAG}_weak.err
python -c "
import json
d=json.loads(open('$O/bench_${TAG}_weak.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d['gather']['by_rank']); print(d['e2e']['value'], d['e2e']['ms_per_step'])" 2>&1 | cut -c1-1500 | tee -a $O/round_$TAG.log
tail -2 $O/bench_${TAG}_weak.err | tee -a $O/round_$TAG.log
echo "=== N=1 same box" | tee -a $O/round_$TAG.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 $FLAGS > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}_n1.err
python -c "
import json
d=json.loads(open('$O/bench_${TAG}_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d['e2e']['value'], d['e2e']['ms_per_step'])" 2>&1 | cut -c1-1500 | tee -a $O/round_$TAG.log
