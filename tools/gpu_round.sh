#!/bin/bash
# One gpurun call of the development loop: GPU tests, A/B probe of library variants, ncu captures.
# usage: tools/gpu_round.sh <tag> [variants...]   (outputs under gpurun_out/<tag>_*)
set -u
TAG=$1; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.txt
cat gpurun_out/${TAG}_pytest.txt
python tools/decode_probe.py 29 > gpurun_out/${TAG}_probe_new.txt 2>&1
for v in "$@"; do
  ALPB200_LIB=$v python tools/decode_probe.py 29 > gpurun_out/${TAG}_probe_$(basename $v .so).txt 2>&1
done
cat gpurun_out/${TAG}_probe_*.txt
