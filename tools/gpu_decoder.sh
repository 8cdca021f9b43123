#!/bin/bash
# GPU visit for the mask decoder: parity tests, timing, ncu launch list.
TAG=${1:-d}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest -q -m gpu tests/test_gpu_decoder.py 2>&1 | tail -15 | tee $O/dec_tests_$TAG.log
timeout 600 python tools/decoder_bench.py --batch 32 --cascade > $O/decoder_bench_$TAG.json 2> $O/decoder_bench_$TAG.err
echo "bench exit $?"; tail -3 $O/decoder_bench_$TAG.err; cat $O/decoder_bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/ncu_decoder_launches_$TAG.csv \
  python tools/decoder_bench.py --batch 32 --steps 1 --warmup 1 > $O/ncu_decoder_$TAG.log 2>&1
echo "ncu exit $?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('$O/ncu_decoder_launches_$TAG.csv')) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
data = rows[hdr + 1:]
half = data[len(data) // 2:]          # second forward (steady state)
agg = collections.OrderedDict()
for r in half:
    v = float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1e-3)
    k = r[ki].split('(')[0][:60]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} {a[0]:4d} {a[1]:10.1f} us {100*a[1]/tot:5.1f}%")
print("total", round(tot, 1), "us for", len(half), "launches (batch 32)")
PY
