#!/bin/bash
# attention probes only
tag=${1:-r2q}
out=gpurun_out/round_$tag.log
mkdir -p gpurun_out
{
for f in $(ls probes/*.so); do
  timeout 120 python tools/attn_probe.py $f 32 2>&1 | tail -1
done
} > $out 2>&1
cat $out
