#!/bin/bash
# compute-sanitizer over every kernel family at small shapes (tools/sanitize_target.py); logs + a summary in gpurun_out/.
#   tools/gpu_sanitize.sh [tag] [tools...]      default tools: memcheck racecheck synccheck
# memcheck: out-of-bounds / misaligned global, shared, tensor-memory accesses; racecheck: shared-memory hazards between
# the warp-specialised roles (mbarrier rings, cluster DSMEM publication, in-place epilogues); synccheck: barrier misuse.
TAG=${1:-r2}
shift
TOOLS=${@:-memcheck racecheck synccheck}
O=gpurun_out
mkdir -p $O
SUM=$O/sanitize_$TAG.summary
: > $SUM
for T in $TOOLS; do
  L=$O/sanitize_${TAG}_$T.log
  EXTRA=""
  [ "$T" == "racecheck" ] && EXTRA="--racecheck-report all"
  timeout 1500 compute-sanitizer --tool $T $EXTRA --print-limit 30 --error-exitcode 7 \
      python tools/sanitize_target.py > $L 2>&1
  RC=$?
  echo "== $T: exit $RC" | tee -a $SUM
  grep -E "^\[sanitize\]|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Misaligned|Barrier error|Traceback" $L | sort | uniq -c | sort -rn | head -25 | tee -a $SUM
done
