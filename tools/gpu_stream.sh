#!/bin/bash
# A/B of the streaming encoder against the block encoder with hard time limits: tools/gpu_stream.sh <tag> [variant .so ...]
set -u
TAG=$1; shift
mkdir -p gpurun_out
export ALPB200_ENCODE_KERNEL=stream
timeout 180 python -m pytest tests/test_gpu_columns.py -m gpu -x -q -k "golden or every_bit_width or appending or overflow_boundaries or rd_every or reference_state" 2>&1 | tail -25
RC=${PIPESTATUS[0]}
if [ "$RC" != "0" ]; then echo "STREAM SMOKE FAILED rc=$RC"; exit 1; fi
for v in alp_b200/libalp_b200.so "$@"; do
  timeout 120 env ALPB200_LIB=$v python tools/probe_enc.py ${LG:-29} 2>&1 | tee -a gpurun_out/${TAG}_enc.txt
done
ALPB200_ENCODE_KERNEL=block timeout 120 python tools/probe_enc.py ${LG:-29} 2>&1 | tee -a gpurun_out/${TAG}_enc.txt
if [ "${NCU:-1}" = "1" ]; then
  timeout 300 tools/ncu_one.sh ${TAG}_encs64 encode_stream_kernel 1 encode 2 27
fi
