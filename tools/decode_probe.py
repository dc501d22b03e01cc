#!/usr/bin/env python
"""Decode throughput on several columns (development probe; run once per library variant via ALPB200_LIB).

    ALPB200_LIB=variants/libalp_b200_cs.so python tools/decode_probe.py [log2_values]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alp_b200  # noqa: E402
from alp_b200 import _abi  # noqa: E402


def _peak():
    import json

    try:
        return float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


PEAK = _peak()  # GB/s, the measured copy bandwidth the roofline fractions refer to


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 29
    n = 1 << lg
    dev = torch.device("cuda:0")
    cols = {}
    cols["decimal_f64 (config 2)"] = alp_b200.generate(n, 2, dev)
    g = torch.Generator(device=dev).manual_seed(1)
    cols["integers<2^20 f64 (no exceptions)"] = torch.randint(0, 1 << 20, (n,), device=dev, generator=g).double()
    cols["2-decimal<100 f64 (bw 14)"] = torch.randint(0, 10000, (n,), device=dev, generator=g).double() / 100.0
    # 0..3 decimals changing from VECTOR to vector: every row-group keeps 4 candidate (e,f) pairs, so each vector runs
    # the second-level sampling (encoder.hpp:241-305) over them
    kk = torch.randint(0, 1000000, (n,), device=dev, generator=g).double()
    dd = ((torch.arange(n, device=dev) // 1024) % 5) % 4  # (period 5: the row-group sampler looks at every 12th vector)
    cols["decimals varying per vector (k=4)"] = kk / torch.tensor([1.0, 10.0, 100.0, 1000.0], dtype=torch.float64, device=dev)[dd]
    del kk, dd
    cols["highprec_f64 (config 3, ALP_RD)"] = alp_b200.generate(n, 3, dev)
    cols["mixed_f32 (config 4)"] = alp_b200.generate(n, 4, dev)
    print("lib", alp_b200.LIB_PATH, " values 2^%d  HBM peak %.0f GB/s (MEASURED_PEAKS.json)" % (lg, PEAK))
    for name, x in cols.items():
        vb = x.element_size()
        col = alp_b200.encode(x)
        pb, ne = col.read_totals()
        out = torch.empty_like(x)
        alp_b200.decode(col, out=out)
        ok = torch.equal(out.view(torch.int64 if vb == 8 else torch.int32), x.view(torch.int64 if vb == 8 else torch.int32))
        meta = col.meta.cpu().numpy().view(_abi.VEC_META_DTYPE).reshape(-1)
        hdr = 5 + vb
        read = pb + hdr * col.n_vectors + (vb + 2) * ne
        algo = read + n * vb
        ms = timed(lambda: alp_b200.decode(col, out=out))
        enc_ws = torch.empty(max(256, alp_b200.lib.alpb200_encode_workspace_bytes(col.n_vectors)), dtype=torch.uint8, device=dev)
        st = alp_b200.rowgroup_init(x)
        ums = timed(lambda: alp_b200.encode(x, st, col=col, workspace=enc_ws, ordered=False), 3) if hasattr(alp_b200.lib, "alpb200_encode_unordered_f64") else float("nan")
        ems = timed(lambda: alp_b200.encode(x, st, col=col, workspace=enc_ws), 3)
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        sms = timed(lambda: alp_b200.decode_sum(col, out=acc))
        print("%-36s ok=%s bits/val=%5.2f exc/vec=%6.1f bw=%s | decode %.3f ms %6.0f GB/s out, %6.0f GB/s algo (%.3f of peak) | encode %.3f ms %6.0f GB/s in, %6.0f algo (%.3f); unordered %.3f ms (%.3f) | sum-scan %.3f ms %6.0f GB/s decoded-equivalent %5.0f GB/s read (%.3f)" % (
            name, ok, 8.0 * read / n, ne / col.n_vectors, sorted(set(meta["bw"].tolist()))[:4], ms, n * vb / ms / 1e6, algo / ms / 1e6, algo / ms / 1e6 / PEAK,
            ems, n * vb / ems / 1e6, algo / ems / 1e6, algo / ems / 1e6 / PEAK, ums, algo / ums / 1e6 / PEAK, sms, n * vb / sms / 1e6, read / sms / 1e6, read / sms / 1e6 / PEAK))
        del col, out


if __name__ == "__main__":
    main()
