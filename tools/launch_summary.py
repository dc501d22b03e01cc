#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and share of the time.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file launches.csv python bench.py --steps 2 --warmup 3 --no-e2e
    python tools/launch_summary.py launches.csv "<command that was profiled>"
"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    tot[name][0] += 1
    tot[name][1] += float(r[14]) / 1e6
total = sum(v[1] for v in tot.values())
print("launch list of `%s` under ncu (cold cache, serialised): %d launches, total %.1f ms" % (sys.argv[2] if len(sys.argv) > 2 else "?", len(rows), total))
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("  %-62s n=%3d %9.3f ms %6.1f%%  (%.3f ms each)" % (name[:62], n, ms, 100.0 * ms / total, ms / n))
