#!/usr/bin/env python
"""SUM-scan timing only (development A/B probe): ALPB200_LIB=variant python tools/probe_scan.py   (KINDS=2,3,4,dec2)"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alp_b200
from tools.probe_enc import timed, column  # noqa: E402
if __name__ == "__main__":
    dev = torch.device("cuda:0")
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6515.0) if os.path.exists("MEASURED_PEAKS.json") else 6515.0
    print("lib", alp_b200.LIB_PATH)
    for kind in os.environ.get("KINDS", "2,3,4,dec2").split(","):
        n = 1 << (28 if kind == "4" else 29)
        x = column(kind, n, dev)
        col = alp_b200.encode(x)
        pb, ne = col.read_totals()
        vb = x.element_size()
        read = pb + ne * (vb + 2) + (n // 1024) * (13 if vb == 8 else 9)
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        ms = timed(lambda: alp_b200.decode_sum(col, out=acc), 10)
        msd = timed(lambda: alp_b200.decode_sum(col, out=acc, flags=alp_b200.SUM_DECIMAL), 10)
        print("kind %-5s sum %.4f ms (%.3f of the read roofline)   decimal %.4f ms (%.3f)" % (kind, ms, read / ms / 1e6 / peak, msd, read / msd / 1e6 / peak))
        del x, col
