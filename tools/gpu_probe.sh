#!/bin/bash
# usage: tools/gpu_probe.sh <tag> [variant .so ...] — decode_probe for the in-tree library and the given variants
set -u
TAG=$1; shift
mkdir -p gpurun_out
python tools/decode_probe.py 29 > gpurun_out/${TAG}_probe.txt 2>&1
for v in "$@"; do
  ALPB200_LIB=$v python tools/decode_probe.py 29 > gpurun_out/${TAG}_probe_$(basename $v .so).txt 2>&1
done
cat gpurun_out/${TAG}_probe*.txt
