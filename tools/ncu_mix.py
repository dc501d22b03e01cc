#!/usr/bin/env python
"""Dynamic SASS opcode mix from an `ncu --page source --csv` export: python tools/ncu_mix.py src.csv n_vectors"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
nvec = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[idx["# Samples"]].strip().isdigit()]
seen, uniq = set(), []
for r in body:
    if r[idx["Address"]] in seen:
        continue
    seen.add(r[idx["Address"]])
    uniq.append(r)
mix, tot = collections.Counter(), 0
for r in uniq:
    n = int(r[idx["Instructions Executed"]] or 0)
    src = re.sub(r"^@!?U?P\d+\s+", "", r[idx["Source"]].strip())
    op = (src.split()[0] if src else "?").split(".")[0]
    mix[op] += n
    tot += n
print("total warp-inst %d  per vector %.1f  per value %.2f" % (tot, tot / nvec, tot / nvec / 32))
for op, n in mix.most_common(28):
    print("%-10s %11d %5.1f%%  %7.1f/vec" % (op, n, 100.0 * n / tot, n / nvec))
