#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_columns.py tests/test_gpu_hardening.py tests/test_gpu_vectors.py -m gpu -x -q 2>&1 | tail -4
for v in alp_b200/libalp_b200.so "$@" alp_b200/libalp_b200.so; do
  ALPB200_LIB=$v timeout 120 python tools/probe_dec.py 28 2>&1 | tee -a gpurun_out/r2ab_dec.txt
done
