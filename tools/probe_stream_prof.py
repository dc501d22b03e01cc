#!/usr/bin/env python
"""Per-phase cycle shares of encode_stream_kernel (profile build):
    ALPB200_LIB=variants/libalp_b200_prof.so python tools/probe_stream_prof.py [log2n]   (KINDS=2,int,3)"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ALPB200_ENCODE_KERNEL"] = "stream"
import alp_b200
NAMES = {0: "c:loop", 1: "c:tile wait", 2: "c:analysis", 3: "c:report", 4: "c:turn wait", 5: "c:space+alloc", 6: "c:exc+pack", 8: "p:exists", 9: "p:lookback", 10: "p:sized wait", 11: "p:done wait", 12: "p:emit", 13: "p:wait_read", 14: "p:retire order"}
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 27
dev = torch.device("cuda:0")
raw = ctypes.CDLL(alp_b200.LIB_PATH)
buf = (ctypes.c_ulonglong * 32)()
for kind in os.environ.get("KINDS", "2,int,3").split(","):
    n = 1 << lg
    if kind == "int":
        g = torch.Generator(device=dev).manual_seed(1)
        x = torch.randint(0, 1 << 20, (n,), device=dev, generator=g).double()
    else:
        x = alp_b200.generate(n, int(kind), dev)
    st = alp_b200.rowgroup_init(x)
    col = alp_b200.DeviceColumn(n // 1024, x.element_size(), dev)
    ws = torch.empty(max(256, alp_b200.lib.alpb200_encode_workspace_bytes(n // 1024)), dtype=torch.uint8, device=dev)
    alp_b200.encode(x, st, col=col, workspace=ws); torch.cuda.synchronize()
    raw.alpb200_debug_stream_profile(buf)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); alp_b200.encode(x, st, col=col, workspace=ws); b.record(); torch.cuda.synchronize()
    raw.alpb200_debug_stream_profile(buf)
    ms = a.elapsed_time(b)
    cyc = ms * 1e-3 * 1.965e9
    print("kind %s: %.3f ms = %.0f cycles; per-role shares of (warps x kernel cycles):" % (kind, ms, cyc))
    comp = sum(buf[i] for i in range(0, 7)); plac = sum(buf[i] for i in range(8, 15))
    for i in sorted(NAMES):
        tot = comp if i < 8 else plac
        print("   %-16s %6.1f%%   (%.1f warp-equivalents busy)" % (NAMES[i], 100.0 * buf[i] / max(tot, 1), buf[i] / cyc / 148))
