#!/bin/bash
# Round-2 record run: GPU tests, both bench arms, launch list, ncu captures of the three hot kernels, probes.
# usage: tools/gpu_r2_full.sh <tag>   (outputs under gpurun_out/<tag>_*)
set -u
TAG=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tail -1 > gpurun_out/${TAG}_env.txt
nproc >> gpurun_out/${TAG}_env.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.txt
cat gpurun_out/${TAG}_pytest.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --cpu-seconds 0.5 > gpurun_out/${TAG}_launches_run.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv "python bench.py --steps 2 --warmup 3 --no-e2e" > gpurun_out/${TAG}_launches_summary.txt 2>&1
timeout 300 tools/ncu_one.sh ${TAG}_dec64 decode_kernel 1 decode 2 30
timeout 300 tools/ncu_one.sh ${TAG}_enc64 encode_kernel 1 encode 2 27
timeout 300 tools/ncu_one.sh ${TAG}_sum64 decode_sum_kernel 1 sum 2 30
timeout 300 python tools/decode_probe.py 29 > gpurun_out/${TAG}_probe.txt 2>&1
cat gpurun_out/${TAG}_probe.txt
