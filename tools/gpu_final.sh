#!/bin/bash
# Final record of a round: the full record run, then compute-sanitizer over every kernel (both vector-order encoders).
set -u
TAG=$1
tools/gpu_r2_full.sh $TAG
OUT=gpurun_out/${TAG}_sanitizer.txt
: > $OUT
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_smoke.py" >> $OUT
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_smoke.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize smoke OK|Error|error|hazard" | head -12 >> $OUT
done
echo "== ALPB200_ENCODE_KERNEL=stream compute-sanitizer --tool memcheck python tools/sanitize_smoke.py" >> $OUT
ALPB200_ENCODE_KERNEL=stream timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_smoke.py 2>&1 | grep -E "ERROR SUMMARY|sanitize smoke OK|Error|error" | head -12 >> $OUT
cat $OUT
