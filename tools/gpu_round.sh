#!/bin/bash
# One GPU-box visit: encoder parity tests, parity report, bench line, ncu launch list, ncu full capture of
# the GEMM kernel.  Everything lands in gpurun_out/ with the tag given as $1.
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
fault() { grep -qiE "illegal memory|launch failure|unspecified launch|misaligned address|mbarrier timeout|CUDA error" "$1"; }
echo "=== tests" | tee $O/round_$TAG.log
timeout 900 python -m pytest -q -m gpu tests/ 2>&1 | tail -15 | tee -a $O/round_$TAG.log
if fault $O/round_$TAG.log; then echo "fault in tests, stopping"; exit 3; fi
echo "=== parity" | tee -a $O/round_$TAG.log
timeout 600 python tests/parity_report.py $O/parity_$TAG.json 2>&1 | tail -8 | tee -a $O/round_$TAG.log
echo "=== smoke" | tee -a $O/round_$TAG.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee -a $O/round_$TAG.log
echo "=== bench" | tee -a $O/round_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
echo "bench exit $?" | tee -a $O/round_$TAG.log
tail -c 3000 $O/bench_$TAG.json | tee -a $O/round_$TAG.log
tail -5 $O/bench_$TAG.err | tee -a $O/round_$TAG.log
echo "=== mask decoder (8 f1)" | tee -a $O/round_$TAG.log
timeout 600 python tools/decoder_bench.py --batch 32 --cascade > $O/decoder_bench_$TAG.json 2> $O/decoder_bench_$TAG.err
echo "decoder bench exit $?" | tee -a $O/round_$TAG.log
cut -c1-260 $O/decoder_bench_$TAG.json | tee -a $O/round_$TAG.log
timeout 600 python tests/cpu_baselines.py $O/cpu_baselines_$TAG.json 2>&1 | tail -3 | tee -a $O/round_$TAG.log
echo "=== CLIP tower (8 f2), cascade" | tee -a $O/round_$TAG.log
timeout 600 python tools/tower_bench.py --batch 32 > $O/tower_bench_$TAG.json 2> $O/tower_bench_$TAG.err
echo "tower bench exit $?" | tee -a $O/round_$TAG.log
cut -c1-260 $O/tower_bench_$TAG.json | tee -a $O/round_$TAG.log
timeout 600 python tools/cascade_bench.py --batch 32 > $O/cascade_bench_$TAG.json 2> $O/cascade_bench_$TAG.err
echo "cascade bench exit $?" | tee -a $O/round_$TAG.log
cut -c1-420 $O/cascade_bench_$TAG.json | tee -a $O/round_$TAG.log
if [ "$2" == "noncu" ]; then exit 0; fi
TOW="python tools/tower_bench.py --batch 32 --steps 1 --warmup 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ncu_tower_launches_$TAG.csv \
    $TOW > $O/ncu_tower_$TAG.log 2>&1
echo "ncu tower list exit $?" | tee -a $O/round_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_global_tc2_kernel -s 24 -c 1 \
    -o $O/prof_attn_plain_$TAG -f $TOW > $O/ncu_full_ap_$TAG.log 2>&1
echo "ncu full plain attention exit $?" | tee -a $O/round_$TAG.log
DEC="python tools/decoder_bench.py --batch 32 --steps 1 --warmup 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/ncu_decoder_launches_$TAG.csv \
    $DEC > $O/ncu_decoder_$TAG.log 2>&1
echo "ncu decoder list exit $?" | tee -a $O/round_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dec_token_layer_kernel -s 2 -c 1 \
    -o $O/prof_dec_token_$TAG -f $DEC > $O/ncu_full_dt_$TAG.log 2>&1
echo "ncu full decoder token exit $?" | tee -a $O/round_$TAG.log
KREG='regex:gemm_bf16_tcgen05_kernel|attention_kernel|attn_global_tc2_kernel|attn_window_tc2_kernel|layernorm_kernel|ln_finalize_kernel|im2col|cast_bf16_kernel|cast_pad_rows_kernel|adapter_embed_kernel|layernorm_nchw_kernel'
BENCH="python bench.py --steps 1 --warmup 3 --batch 32 --skip-cpu-baseline --skip-e2e --skip-decoder"
echo "=== ncu launch list" | tee -a $O/round_$TAG.log
$BENCH > $O/plain_$TAG.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 729 -c 243 --csv \
    --log-file $O/launches_$TAG.csv $BENCH > $O/ncu_list_$TAG.log 2>&1
echo "ncu list exit $?" | tee -a $O/round_$TAG.log
echo "=== ncu full (gemm: qkv / proj / lin1 / lin2 of one block)" | tee -a $O/round_$TAG.log
$BENCH > $O/plain2_$TAG.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_kernel -s 461 -c 4 \
    -o $O/prof_gemm_$TAG -f $BENCH > $O/ncu_full_$TAG.log 2>&1
echo "ncu full gemm exit $?" | tee -a $O/round_$TAG.log
echo "=== ncu full (attention)" | tee -a $O/round_$TAG.log
$BENCH > $O/plain3_$TAG.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:attn_global_tc2_kernel -s 13 -c 1 \
    -o $O/prof_attn_global_$TAG -f $BENCH > $O/ncu_full_ag_$TAG.log 2>&1
echo "ncu full attn global exit $?" | tee -a $O/round_$TAG.log
$BENCH > $O/plain4_$TAG.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:attn_window_tc2_kernel -s 90 -c 1 \
    -o $O/prof_attn_window_$TAG -f $BENCH > $O/ncu_full_aw_$TAG.log 2>&1
echo "ncu full attn window exit $?" | tee -a $O/round_$TAG.log
tail -3 $O/ncu_full_aw_$TAG.log | tee -a $O/round_$TAG.log
