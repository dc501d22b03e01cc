#!/usr/bin/env python
"""Run ONE hot-path launch a few times, for profiling under ncu (never a bench value).

    ncu --set full --import-source on --clock-control none -k regex:encode_kernel -s 1 -c 1 -o gpurun_out/enc \
        python tools/ncu_probe.py encode 2 27          # what (encode|encode_unordered|decode|sum|minmax|filter), column kind (2|3|4|int), log2(values)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alp_b200  # noqa: E402


def main():
    what, kind, lg = sys.argv[1], sys.argv[2], int(sys.argv[3])
    n = 1 << lg
    dev = torch.device("cuda:0")
    if kind == "int":
        g = torch.Generator(device=dev).manual_seed(1)
        x = torch.randint(0, 1 << 20, (n,), device=dev, generator=g).double()
    elif kind == "k4":  # 0..3 decimals changing from vector to vector: second-level sampling over 4 candidate (e,f) pairs
        g = torch.Generator(device=dev).manual_seed(1)
        kk = torch.randint(0, 1000000, (n,), device=dev, generator=g).double()
        dd = ((torch.arange(n, device=dev) // 1024) % 5) % 4
        x = kk / torch.tensor([1.0, 10.0, 100.0, 1000.0], dtype=torch.float64, device=dev)[dd]
    else:
        x = alp_b200.generate(n, int(kind), dev)
    st = alp_b200.rowgroup_init(x)
    col = alp_b200.DeviceColumn(n // 1024, x.element_size(), dev)
    ws = torch.empty(max(256, alp_b200.lib.alpb200_encode_workspace_bytes(n // 1024)), dtype=torch.uint8, device=dev)
    out = torch.empty_like(x)
    acc = torch.zeros(1, dtype=torch.float64, device=dev)
    ordered = what != "encode_unordered"
    alp_b200.encode(x, st, col=col, workspace=ws, ordered=ordered)
    col.read_totals()  # also learns the widest block: the decoders size their stages from it
    for _ in range(3):
        if what.startswith("encode"):
            alp_b200.encode(x, st, col=col, workspace=ws, ordered=ordered)
        if what == "decode":
            alp_b200.decode(col, out=out)
        if what == "sum":
            alp_b200.decode_sum(col, out=acc)
        if what == "minmax":
            alp_b200.decode_minmax(col)
        if what == "filter":
            alp_b200.decode_filter(col, "<", 500.0)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
