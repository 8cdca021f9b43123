#!/bin/bash
# Round-2 first GPU visit: full GPU test suite (new batch-32 parity tests), a bench line with the new legs, sanitizer.
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
echo "=== tests" | tee $O/round_$TAG.log
timeout 1800 python -m pytest -q -m gpu tests/ -s 2>&1 | grep -vE "^\s*$" | tail -40 | tee -a $O/round_$TAG.log
echo "=== bench" | tee -a $O/round_$TAG.log
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
echo "bench exit $?" | tee -a $O/round_$TAG.log
tail -c 4000 $O/bench_$TAG.json | tee -a $O/round_$TAG.log
tail -5 $O/bench_$TAG.err | tee -a $O/round_$TAG.log
echo "=== reference arm" | tee -a $O/round_$TAG.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err
cut -c1-300 $O/bench_ref_$TAG.json | tee -a $O/round_$TAG.log
echo "=== sanitizer" | tee -a $O/round_$TAG.log
timeout 2400 tools/gpu_sanitize.sh $TAG 2>&1 | tail -60 | tee -a $O/round_$TAG.log
