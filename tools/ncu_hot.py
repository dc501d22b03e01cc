#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: stall totals by reason and the hottest SASS lines.

    ncu -i prof.ncu-rep --page source --csv -k regex:encode_kernel > src.csv ; python tools/ncu_hot.py src.csv [N]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr) and r[idx["# Samples"]].strip().isdigit()]
    stall_cols = [h for h in hdr if h.startswith("stall_") and not h.endswith("(Not Issued)")]
    tot = {h: 0 for h in stall_cols}
    samples = 0
    for r in body:
        samples += int(r[idx["# Samples"]] or 0)
        for h in stall_cols:
            tot[h] += int(r[idx[h]] or 0)
    print("samples", samples, " instructions(SASS lines)", len(body))
    for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
        print("  %-24s %8d  %5.1f%%" % (h, v, 100.0 * v / max(samples, 1)))
    print("hottest lines:")
    body.sort(key=lambda r: -int(r[idx["# Samples"]] or 0))
    for r in body[:top]:
        reasons = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
        print("  %6s  %-70s %s" % (r[idx["# Samples"]], r[idx["Source"]][:70], " ".join("%s:%d" % (n, v) for v, n in reasons if v)))


if __name__ == "__main__":
    main()
