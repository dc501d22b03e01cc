#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref) in this container.

The GPU box has no /root/reference, so everything the parity tests need from it is frozen here:

* ``reference_vectors.npz`` — every single-vector case of the reference's own test
  (test/test_alp_sample.cpp:191-227): the first 1024 values of each column named by the descriptor tables
  in data/include/*.hpp, the two golden asserts of that test (``bit_width``, ``exceptions_count``;
  test_alp_sample.cpp:178-179) and the complete output of the reference primitives on that vector
  (row-group state, e, f, bw, base, exceptions, positions, packed words; for ALP_RD the cut, dictionary,
  both packed streams and the left-part exceptions).
* ``reference_columns.npz`` — three multi-row-group samples (data/1_rg_data_sample/*.bin, 131072 f64 = 128
  vectors = 2 row-groups each) with the reference-encoded column container, and the three synthetic
  columns of SURVEY.md §8d (4 row-groups each) with the per-vector metadata the reference produces.

Usage:  python tools/make_golden.py        (needs /root/reference and a built oracle/_ref)
"""
import hashlib
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

ENTRY = re.compile(
    r'\{\s*(\d+),\s*"([^"]+)",\s*[^"{}]*"([^"]+\.csv)",\s*[^"{}]*"([^"]*)",\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\s*(?:,\s*(true|false))?\s*\}'
)


def descriptors(header, func, csv_dir):
    """Parse one ALPColumnDescriptor table (data/include/column.hpp:30-40) out of a header."""
    text = open(os.path.join(REF, "data", "include", header)).read()
    start = text.index(func)
    end = text.index("return", start)
    rows = []
    for m in ENTRY.finditer(text[start:end]):
        csv = m.group(3)
        path = os.path.join(REF, csv.lstrip("/")) if csv.startswith("/data") else os.path.join(REF, "data", csv_dir, csv)
        rows.append(dict(name=m.group(2), csv=path, golden_exceptions=int(m.group(7)), golden_bw=int(m.group(8))))
    return rows


def load_csv(path, dtype):
    vals = [float(tok.rstrip(",")) for tok in open(path).read().split()]  # std::stod / std::stof stop at the comma
    if dtype == np.float32:
        return np.array(vals, dtype=np.float64).astype(np.float32)[:1024]
    return np.array(vals, dtype=np.float64)[:1024]


def vector_cases():
    groups = [
        ("alp_dataset", np.float64, descriptors("double/alp_dataset.hpp", "get_alp_dataset()", "samples")),
        ("generated", np.float64, descriptors("generated_columns.hpp", "get_generated_cols()", "generated")),
        ("edge_case", np.float64, descriptors("edge_case.hpp", "get_edge_case()", "edge_case")),
        ("double_test", np.float64, descriptors("double/alp_dataset.hpp", "get_double_test_dataset()", "double")),
        ("float_test", np.float32, descriptors("float/test.hpp", "get_float_test_dataset()", "samples")),
        ("float_edge_case", np.float32, descriptors("float/edge_case.hpp", "get_float_edge_case()", "edge_case")),
    ]
    expect = {"alp_dataset": 30, "generated": 65, "edge_case": 1, "double_test": 1, "float_test": 5, "float_edge_case": 1}
    for g, _, rows in groups:
        assert len(rows) == expect[g], (g, len(rows))
    return groups


def main():
    R = po.reference()
    assert R is not None, "build oracle/_ref first (make -C oracle ref)"
    os.makedirs(OUT, exist_ok=True)
    arrays, index = {}, []
    for group, dtype, rows in vector_cases():
        for row in rows:
            x = load_csv(row["csv"], dtype)
            assert x.shape[0] == 1024, row
            st = R.init(x)
            cid = "c%03d" % len(index)
            arrays[cid + "_input"] = x
            arrays[cid + "_state"] = st
            entry = dict(
                id=cid,
                group=group,
                name=row["name"],
                file=os.path.basename(row["csv"]),
                relpath=os.path.relpath(row["csv"], REF),  # where the reference's own test looks for it (under ALP_CMAKE_SOURCE_DIR)
                dtype=np.dtype(dtype).name,
                golden_bw=row["golden_bw"],
                golden_exceptions=row["golden_exceptions"],
                scheme=int(st["scheme"][0]),
            )
            ut = np.uint64 if dtype == np.float64 else np.uint32
            if entry["scheme"] == 2:
                r = R.encode(x, st)
                bw, base = R.analyze_ffor(r["enc"])
                packed = R.ffor(r["enc"].view(ut), bw, int(base))
                dec = R.patch(R.decode(R.unffor(packed, bw, int(base), ut).view(r["enc"].dtype), r["f"], r["e"]), r["exc"], r["pos"])
                assert dec.tobytes() == x.tobytes(), row
                assert bw == row["golden_bw"] and r["cnt"] == row["golden_exceptions"], (row, bw, r["cnt"])
                entry.update(e=r["e"], f=r["f"], bw=bw, base=int(base), cnt=r["cnt"])
                arrays[cid + "_enc"] = r["enc"]
                arrays[cid + "_exc"] = r["exc"]
                arrays[cid + "_pos"] = r["pos"]
                arrays[cid + "_packed"] = packed
            else:
                r = R.rd_encode(x, st)
                rbw, lbw = int(st["right_bw"][0]), int(st["left_bw"][0])
                pr = R.ffor(r["right"], rbw, 0)
                pl = R.ffor(r["left"], lbw, 0)
                dec = R.rd_decode(R.unffor(pr, rbw, 0, ut), R.unffor(pl, lbw, 0, np.uint16), r["exc"], r["pos"], st)
                assert dec.tobytes() == x.tobytes(), row
                entry.update(right_bw=rbw, left_bw=lbw, dict_size=int(st["dict_size"][0]), cnt=r["cnt"])
                arrays[cid + "_right"] = r["right"]
                arrays[cid + "_left"] = r["left"]
                arrays[cid + "_exc"] = r["exc"]
                arrays[cid + "_pos"] = r["pos"]
                arrays[cid + "_packed_right"] = pr
                arrays[cid + "_packed_left"] = pl
            index.append(entry)
    arrays["index_json"] = np.frombuffer(json.dumps(index).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "reference_vectors.npz"), **arrays)
    n_rd = sum(1 for e in index if e["scheme"] == 1)
    print("reference_vectors.npz: %d cases (%d ALP with golden bw/exceptions asserted, %d ALP_RD)" % (len(index), len(index) - n_rd, n_rd))

    # ---- multi-row-group columns ----
    cols, cindex = {}, []

    def add_column(name, x, keep_input):
        col = R.encode_column(x, n_threads=8)
        dec = R.decode_column(col, n_threads=8)
        assert dec.tobytes() == x.tobytes(), name
        t = col.trimmed()
        states = np.concatenate([R.init(x, off) for off in range(0, x.shape[0], 102400)])
        if keep_input:
            cols[name + "_input"] = x
        cols[name + "_meta"] = t.meta
        cols[name + "_states"] = states
        packed = t.packed[: t.packed_bytes].copy()
        exc_val = t.exc_val[: t.n_exceptions].copy()
        exc_pos = t.exc_pos[: t.n_exceptions].copy()
        if keep_input:  # real data: keep the whole container; synthetic: digests only (the payload is random bits)
            cols[name + "_packed"] = packed
            cols[name + "_exc_val"] = exc_val
            cols[name + "_exc_pos"] = exc_pos
        cols[name + "_totals"] = t.totals.copy()
        bits = 8.0 * t.compressed_bytes() / x.shape[0]
        sha = {k: hashlib.sha256(v.tobytes()).hexdigest() for k, v in (("packed", packed), ("exc_val", exc_val), ("exc_pos", exc_pos))}
        cindex.append(dict(name=name, dtype=x.dtype.name, n_values=int(x.shape[0]), has_input=bool(keep_input), bits_per_value=bits, sha256=sha))
        print("  %-28s %8d values  %6.2f bits/value  schemes=%s" % (name, x.shape[0], bits, sorted(set(t.meta["scheme"].tolist()))))

    for name in ("city_temperature_f_tw", "food_prices_tw", "gov26_tw"):
        x = np.fromfile(os.path.join(REF, "data", "1_rg_data_sample", name + ".bin"), dtype=np.float64)
        add_column(name, x[: (x.shape[0] // 1024) * 1024], True)
    for kind, name in ((2, "synthetic_decimal_f64"), (3, "synthetic_highprec_f64"), (4, "synthetic_mixed_f32")):
        x = po.generate(4 * 102400, kind)
        add_column(name, x, False)  # the input is regenerated from the seed by the tests
        cindex[-1].update(kind=kind, first_values=[float(v) for v in x[:4]], xor_checksum=int(np.bitwise_xor.reduce(x.view(np.uint64 if kind != 4 else np.uint32))))
    cols["index_json"] = np.frombuffer(json.dumps(cindex).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "reference_columns.npz"), **cols)
    for f in ("reference_vectors.npz", "reference_columns.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
