#!/bin/bash
set -u
mkdir -p gpurun_out
tools/gpu_quick.sh $1 variants/libalp_b200_u32.so variants/libalp_b200_u4.so
for v in alp_b200/libalp_b200.so variants/libalp_b200_u32.so; do
  ALPB200_ENCODE_KERNEL=stream ALPB200_LIB=$v KINDS=2,int timeout 120 python tools/probe_enc.py 29 2>&1 | sed 's/^/stream: /' | tee -a gpurun_out/$1_enc.txt
done
