#!/bin/bash
# Development loop on the GPU box with hard time limits (a hung kernel must not burn the budget):
#   tools/gpu_quick.sh <tag> [variant .so ...]  -> a short parity smoke test, then tools/probe_enc.py per library
set -u
TAG=$1; shift
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_columns.py -m gpu -x -q -k "golden or every_bit_width or completion_order or appending or overflow_boundaries or rd_every" 2>&1 | tail -4
RC=${PIPESTATUS[0]}
if [ "$RC" != "0" ]; then echo "SMOKE FAILED rc=$RC"; exit 1; fi
for v in alp_b200/libalp_b200.so "$@" alp_b200/libalp_b200.so; do
  timeout 90 env ALPB200_LIB=$v python tools/probe_enc.py ${LG:-29} 2>&1 | tee -a gpurun_out/${TAG}_enc.txt
done
