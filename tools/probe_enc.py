#!/usr/bin/env python
"""Encode timing only (development A/B probe): ALPB200_LIB=variant python tools/probe_enc.py [log2n]   (KINDS=2,int,dec2,k4,3,4)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alp_b200
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def column(kind, n, dev):
    g = torch.Generator(device=dev).manual_seed(1)
    if kind == "int":
        return torch.randint(0, 1 << 20, (n,), device=dev, generator=g).double()
    if kind == "dec2":
        return torch.randint(0, 10000, (n,), device=dev, generator=g).double() / 100.0
    if kind == "neg":  # signed 2-decimal values: encoded integers straddle zero
        return (torch.randint(0, 2000000, (n,), device=dev, generator=g).double() - 1000000.0) / 100.0
    if kind == "k4":
        kk = torch.randint(0, 1000000, (n,), device=dev, generator=g).double()
        dd = ((torch.arange(n, device=dev) // 1024) % 5) % 4
        return kk / torch.tensor([1.0, 10.0, 100.0, 1000.0], dtype=torch.float64, device=dev)[dd]
    return alp_b200.generate(n, int(kind), dev)
def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 29
    dev = torch.device("cuda:0")
    print("lib", alp_b200.LIB_PATH)
    for kind in os.environ.get("KINDS", "2,int,dec2,neg,k4,3,4").split(","):
        n = 1 << lg
        x = column(kind, n, dev)
        st = alp_b200.rowgroup_init(x)
        col = alp_b200.DeviceColumn(n // 1024, x.element_size(), dev)
        ws = torch.empty(max(256, alp_b200.lib.alpb200_encode_workspace_bytes(n // 1024)), dtype=torch.uint8, device=dev)
        eo = timed(lambda: alp_b200.encode(x, st, col=col, workspace=ws))
        pb, ne = col.read_totals()
        y = alp_b200.decode(col)
        ib = torch.int64 if x.element_size() == 8 else torch.int32
        ok = torch.equal(x.view(ib), y.view(ib))
        eu = timed(lambda: alp_b200.encode(x, st, col=col, workspace=ws, ordered=False))
        ti = timed(lambda: alp_b200.rowgroup_init(x, states=st))
        print("kind %-5s ok=%s ordered %.4f ms  unordered %.4f ms  init %.4f ms  (%.2f bits/value, %.1f exc/vec)" % (kind, ok, eo, eu, ti, 8.0 * (pb + 10 * ne) / n, ne / (n / 1024)))
        del x, col, y


if __name__ == "__main__":
    main()
