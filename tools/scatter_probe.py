#!/usr/bin/env python
"""Multi-GPU check: rank 0 compresses a column, shards travel over NCCL (NVLink), every rank decodes its shard.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/scatter_probe.py [log2_values]

Prints the scatter time (set-up, outside any decode timing) and the per-rank decode rate; verifies every shard bit for
bit against the regenerated slice of the global column.
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alp_b200  # noqa: E402
from alp_b200 import shard  # noqa: E402


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    n = 1 << lg
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    col = None
    if rank == 0:
        x = alp_b200.generate(n, 2, dev)
        col = shard.column_tensors(alp_b200.encode(x))
        del x
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mine, (first, count) = shard.scatter_column(col, src=0, value_bytes=8, device=dev)
    torch.cuda.synchronize()
    dist.barrier()
    t_scatter = time.perf_counter() - t0
    dcol = shard.tensors_to_device_column(mine, 8, dev)
    out = torch.empty(count * 1024, dtype=torch.float64, device=dev)
    alp_b200.decode(dcol, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        alp_b200.decode(dcol, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    want = alp_b200.generate(count * 1024, 2, dev, first_index=first * 1024)
    ok = torch.equal(out.view(torch.int64), want.view(torch.int64))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    nbytes = sum(t.numel() * t.element_size() for t in mine.values())
    print("rank %d: vectors [%d,+%d) shard %.1f MB decode %.3f ms (%.0f GB/s out) ok=%s" % (rank, first, count, nbytes / 1e6, ms, count * 8192 / ms / 1e6, ok), flush=True)
    if rank == 0:
        print("scatter of 2^%d values over %d ranks: %.1f ms; all shards bit-exact: %s" % (lg, world, t_scatter * 1e3, bool(flag.item())), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
