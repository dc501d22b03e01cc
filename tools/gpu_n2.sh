#!/bin/bash
# 2-GPU check: the NCCL scatter test and the bench line at N = 2 (scatter block, e2e link probe with both ranks copying)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "scatter or shard" 2>&1 | tail -4 | tee gpurun_out/$1_n2_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/$1_bench_n2.json 2> gpurun_out/$1_bench_n2.err
tail -2 gpurun_out/$1_bench_n2.err
python -c "
import json,sys
d=json.load(open('gpurun_out/$1_bench_n2.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('link'), d['e2e'].get('link_frac'), d.get('scatter'))
"
