#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_columns.py -m gpu -x -q -k "minmax or decode_sum or decimal or filter" 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0.5 > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err
tail -3 gpurun_out/r2ac_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2ac_bench.json'))
for k in ('2','3','4'):
    c=d['configs'][k]; print(k, {kk:(round(v['ms'],3), round(v['roofline_frac'],3)) for kk,v in c.items() if isinstance(v,dict) and 'ms' in v})
"
