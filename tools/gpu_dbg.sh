#!/bin/bash
set -u
mkdir -p gpurun_out
ALPB200_LIB=variants/libalp_b200_prof.so timeout 200 python tools/probe_stream_prof.py 28 2>&1 | tee gpurun_out/dbg_prof.txt
