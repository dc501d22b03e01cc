#!/bin/bash
# What the driver runs at round end, in one call: GPU tests, smoke(), both bench arms.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/check_bench_ref.json 2> gpurun_out/check_bench_ref.err; echo "ref rc=$?"
( time timeout 900 python bench.py ) > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/check_bench.err
python -c "
import json
d=json.load(open('gpurun_out/check_bench.json'))
print(d['metric'], round(d['value'],1), d['unit'], 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], d['clocks'])
"
