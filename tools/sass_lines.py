#!/usr/bin/env python
"""Attribute executed warp instructions / stall samples to CUDA source lines.

    python tools/sass_lines.py <object.o> <kernel-substring> <counts.csv from tools/ncu_compact.py> [n_vectors] [top]

Disassembles the kernel from the local object with `nvdisasm --print-line-info` (same build as the one profiled) and
joins on the instruction offset."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def disassemble(obj, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    text = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    lines = text.split("\n")
    out, cur, inside = {}, ("?", 0), False
    for ln in lines:
        if ln.startswith("//---") and ".text." in ln:
            inside = kernel_sub in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m:
            out[int(m.group(1), 16)] = (cur, m.group(2))
    return out


def main():
    obj, ksub, counts = sys.argv[1], sys.argv[2], sys.argv[3]
    nvec = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 50
    dis = disassemble(obj, ksub)
    by_line = collections.defaultdict(lambda: [0, 0])
    tot_n = tot_s = miss = 0
    for ln in open(counts):
        off, n, s, op = ln.strip().split(",")
        off, n, s = int(off, 16), int(n), int(s)
        key = dis.get(off, (("?", 0), op))[0]
        if off not in dis:
            miss += 1
        by_line[key][0] += n
        by_line[key][1] += s
        tot_n += n
        tot_s += s
    print("total warp-inst %d (%.1f per vector), samples %d, unmatched offsets %d" % (tot_n, tot_n / nvec, tot_s, miss))
    src_cache = {}

    def text(key):
        f, l = key
        for d in ("alp_b200/csrc", "include"):
            p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().split("\n")
                return src_cache[p][l - 1].strip()[:100] if 0 < l <= len(src_cache[p]) else ""
        return ""

    print("-- by instructions executed")
    for key, (n, s) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%8.1f/vec %5.1f%%  smp %5.1f%%  %-18s:%-4d %s" % (n / nvec, 100.0 * n / max(tot_n, 1), 100.0 * s / max(tot_s, 1), key[0], key[1], text(key)))
    print("-- by stall samples")
    for key, (n, s) in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:25]:
        print("%8.1f/vec %5.1f%%  smp %5.1f%%  %-18s:%-4d %s" % (n / nvec, 100.0 * n / max(tot_n, 1), 100.0 * s / max(tot_s, 1), key[0], key[1], text(key)))


if __name__ == "__main__":
    main()
