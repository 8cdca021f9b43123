#!/bin/bash
# ncu launch list of the attention forward / backward kernels of one training step (batch 4)
tag=${1:-r2tc}
O=gpurun_out
mkdir -p $O
TRAIN="python tools/train_bench.py --batch 4"
$TRAIN > $O/plain_$tag.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:attn_" -s 96 -c 96 --csv \
    --log-file $O/ncu_attn_train_launches_${tag}_b4.csv $TRAIN > $O/ncu_list_$tag.log 2>&1
echo "ncu exit $?"
python - <<PY
import csv,collections
rows=list(csv.reader(open('$O/ncu_attn_train_launches_${tag}_b4.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; data=rows[hi+1:]
kn=h.index('Kernel Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
tot=collections.Counter(); cnt=collections.Counter()
for r in data:
    if len(r)<=mv: continue
    v=float(r[mv].replace(',',''))*{'ns':1e-6,'us':1e-3,'ms':1.0}.get(r[mu],1e-6)
    n=r[kn].split('(')[0][-48:]
    tot[n]+=v; cnt[n]+=1
for k,v in tot.most_common(): print(f"{k:50s} {cnt[k]:4d} launches  {v/cnt[k]:8.3f} ms each")
PY
