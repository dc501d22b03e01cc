#!/usr/bin/env python
"""Decode / sum-scan timing only (development A/B probe): ALPB200_LIB=variant python tools/probe_dec.py [log2n]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alp_b200
def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 29
dev = torch.device("cuda:0")
print("lib", alp_b200.LIB_PATH)
for kind in [int(k) for k in os.environ.get("KINDS", "2,3,4").split(",")]:
    x = alp_b200.generate(1 << lg, kind, dev)
    col = alp_b200.encode(x); col.read_totals()
    out = torch.empty_like(x); acc = torch.zeros(1, dtype=torch.float64, device=dev)
    for rep in range(2):
        print("kind %d  decode %.4f ms  sum %.4f ms" % (kind, timed(lambda: alp_b200.decode(col, out=out)), timed(lambda: alp_b200.decode_sum(col, out=acc))))
    del x, col, out
