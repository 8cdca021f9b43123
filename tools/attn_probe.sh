#!/bin/bash
# Builds side copies of the library with the bottleneck probes of attention_tc2.cu / attention_tc_win2.cu compiled in
# (no GPU needed), into probes/ (git-ignored, travels with the gpurun snapshot).  tools/attn_probe.py times them.
set -e
cd "$(dirname "$0")/../camouflaged_vlm_b200/csrc"
make -s
mkdir -p ../../probes build_probe
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr"
OTHERS=$(ls build/*.o | grep -v -e attention_tc2.o -e attention_tc_win2.o)
for spec in "g0:0:0" "g1:1:0" "g2:2:0" "g3:3:0" "g4:4:0" "w1:0:1" "w2:0:2" "w3:0:3" "w4:0:4" "w8:0:8" "w10:0:10" "w11:0:11"; do
  IFS=: read name gp wp <<< "$spec"
  nvcc $FLAGS -DG2_PROBE=$gp -c attention_tc2.cu -o build_probe/attention_tc2_$name.o &
  nvcc $FLAGS -DW2_PROBE=$wp -c attention_tc_win2.cu -o build_probe/attention_tc_win2_$name.o &
  wait
  nvcc $ARCH -shared -o ../../probes/libcvlm_$name.so $OTHERS build_probe/attention_tc2_$name.o build_probe/attention_tc_win2_$name.o -cudart static
done
ls -la ../../probes
