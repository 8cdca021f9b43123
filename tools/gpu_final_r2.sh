#!/bin/bash
# Round-2 closing visit: everything the judge reads, in one place.  Results land in gpurun_out/ with the tag $1.
TAG=${1:-r2final}
O=gpurun_out
mkdir -p $O
L=$O/round_$TAG.log
echo "=== tests" | tee $L
timeout 2400 python -m pytest -q -m gpu tests/ -s 2>&1 | grep -vE "^\s*$" | tail -45 | cut -c1-420 | tee -a $L
echo "=== smoke" | tee -a $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee -a $L
echo "=== parity" | tee -a $L
timeout 900 python tests/parity_report.py $O/parity_$TAG.json 2>&1 | tail -3 | cut -c1-1500 | tee -a $L
echo "=== bench" | tee -a $L
timeout 1500 python bench.py --steps 5 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
echo "bench exit $?" | tee -a $L
python -c "
import json
d=json.loads(open('$O/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','clocks','parity','adapter_training','cpu_baseline')})
print(d['roofline']); print(d['e2e']); print(d['kernels']); print(d['torch_gpu_baseline'])" 2>&1 | cut -c1-3500 | tee -a $L
tail -3 $O/bench_$TAG.err | tee -a $L
echo "=== reference arm" | tee -a $L
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err
cut -c1-400 $O/bench_ref_$TAG.json | tee -a $L
echo "=== train step by kind, decoder, tower, cascade" | tee -a $L
timeout 600 python tools/train_bench.py --batch 8 > $O/train_bench_$TAG.json 2> $O/train_bench_$TAG.err
cut -c1-1800 $O/train_bench_$TAG.json | tee -a $L
timeout 600 python tools/decoder_bench.py --batch 32 --cascade > $O/decoder_bench_$TAG.json 2> $O/decoder_bench_$TAG.err
cut -c1-300 $O/decoder_bench_$TAG.json | tee -a $L
timeout 600 python tools/tower_bench.py --batch 32 > $O/tower_bench_$TAG.json 2> $O/tower_bench_$TAG.err
cut -c1-300 $O/tower_bench_$TAG.json | tee -a $L
timeout 600 python tools/cascade_bench.py --batch 32 > $O/cascade_bench_$TAG.json 2> $O/cascade_bench_$TAG.err
cut -c1-420 $O/cascade_bench_$TAG.json | tee -a $L
if [ "$2" == "noncu" ]; then exit 0; fi
KREG='regex:gemm_bf16_tcgen05_kernel|attention_kernel|attn_global_tc2_kernel|attn_window_tc2_kernel|layernorm_kernel|ln_finalize_kernel|im2col|cast_bf16_kernel|cast_pad_rows_kernel|adapter_embed_kernel|layernorm_nchw_kernel'
BENCH="python bench.py --steps 1 --warmup 3 --batch 32 --skip-cpu-baseline --skip-e2e --skip-decoder --skip-train"
echo "=== ncu launch list of one bench step" | tee -a $L
$BENCH > $O/plain_$TAG.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 726 -c 242 --csv \
    --log-file $O/ncu_launches_${TAG}_b32.csv $BENCH > $O/ncu_list_$TAG.log 2>&1
echo "ncu list exit $?" | tee -a $L
echo "=== ncu full: qkv / proj / lin1 / lin2 of one block" | tee -a $L
$BENCH > $O/plain2_$TAG.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_kernel -s 461 -c 4 \
    -o $O/prof_gemm_$TAG -f $BENCH > $O/ncu_full_$TAG.log 2>&1
echo "ncu full gemm exit $?" | tee -a $L
TRAIN="python tools/train_bench.py --batch 4"
echo "=== ncu launch list of the backward kernels (one training step, batch 4)" | tee -a $L
$TRAIN > $O/plain3_$TAG.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:attn_bwd|wgrad|ln_bwd|eg_wgrad|layernorm_train" -s 420 -c 210 --csv \
    --log-file $O/ncu_train_launches_${TAG}_b4.csv $TRAIN > $O/ncu_trainlist_$TAG.log 2>&1
echo "ncu train list exit $?" | tee -a $L
echo "=== ncu full: attention backward (64x64 dkv, 14x14 window)" | tee -a $L
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dkv64_kernel -s 8 -c 1 \
    -o $O/prof_attn_bwd_dkv64_$TAG -f $TRAIN > $O/ncu_full_bd_$TAG.log 2>&1
echo "ncu full dkv64 exit $?" | tee -a $L
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_win14_kernel -s 60 -c 1 \
    -o $O/prof_attn_bwd_win14_$TAG -f $TRAIN > $O/ncu_full_bw_$TAG.log 2>&1
echo "ncu full win14 exit $?" | tee -a $L
tail -2 $O/ncu_full_bw_$TAG.log | tee -a $L
