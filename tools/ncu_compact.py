#!/usr/bin/env python
"""ncu `--page source --csv` (SASS view) -> compact `offset,executed,samples,opcode` lines (a few hundred KiB), which
tools/sass_lines.py joins with `nvdisasm --print-line-info` of the same build to attribute work to CUDA source lines."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[idx["# Samples"]].strip().isdigit()]
seen = {}
for r in body:
    a = int(r[idx["Address"]], 16) if r[idx["Address"]].startswith("0x") else int(r[idx["Address"]])
    if a in seen:
        continue
    src = re.sub(r"^@!?U?P\d+\s+", "", r[idx["Source"]].strip())
    seen[a] = (int(r[idx["Instructions Executed"]] or 0), int(r[idx["# Samples"]] or 0), (src.split()[0] if src else "?"))
base = min(seen)
with open(sys.argv[2], "w") as fh:
    for a in sorted(seen):
        n, s, op = seen[a]
        fh.write("%x,%d,%d,%s\n" % (a - base, n, s, op))
