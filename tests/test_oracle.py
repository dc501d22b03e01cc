"""Pins the CPU oracle (oracle/alp_oracle.c, a plain-C restatement of the reference's vector path).

1. Against the committed golden vectors: every known answer the reference's own test holds for this path
   (test/test_alp_sample.cpp:172-179 — bit_width and exceptions_count per fixture column) plus the full output of
   the unmodified reference on those vectors (tests/golden/reference_vectors.npz, tools/make_golden.py).
2. Against the unmodified reference compiled from /root/reference (oracle/_ref), function by function on random
   and adversarial inputs — skipped where oracle/_ref was never built.
"""
import hashlib

import numpy as np
import pytest

from conftest import host_column_from_golden


def _same(a, b):
    return np.asarray(a).tobytes() == np.asarray(b).tobytes()


def _ut(x):
    return np.uint64 if x.dtype.itemsize == 8 else np.uint32


# ---------------------------------------------------------------------------------------------------------------------
# 1. golden vectors
# ---------------------------------------------------------------------------------------------------------------------
def test_config1_known_answer(golden_vectors, port):
    """BASELINE config 1: data/double/test_0.csv (1024 x 10.23) → scheme ALP, k=1, e=17 f=15 bw=0 base=1023, no
    exceptions, and a bit-exact round trip on the CPU."""
    g = golden_vectors
    c = [c for c in g.index if c["group"] == "double_test"][0]
    x = g[c["id"] + "_input"]
    assert np.all(x == 10.23)
    st = port.init(x)
    assert (int(st["scheme"][0]), int(st["k"][0]), st["combos"][0, 0].tolist()) == (2, 1, [17, 15])
    r = port.encode(x, st)
    bw, base = port.analyze_ffor(r["enc"])
    assert (r["e"], r["f"], bw, int(base), r["cnt"]) == (17, 15, 0, 1023, 0)
    dec = port.patch(port.falp(port.ffor(r["enc"].view(np.uint64), bw, int(base)), bw, int(base), r["f"], r["e"]), r["exc"], r["pos"])
    assert _same(dec, x)


def test_port_reproduces_every_golden_alp_case(golden_vectors, port):
    g = golden_vectors
    cases = g.cases(scheme=2)
    assert len(cases) == 98
    for c in cases:
        cid, x = c["id"], g[c["id"] + "_input"]
        st = port.init(x)
        assert _same(st, g[cid + "_state"]), c["name"]
        r = port.encode(x, st)
        bw, base = port.analyze_ffor(r["enc"])
        # the reference test's two golden asserts (test_alp_sample.cpp:178-179)
        assert r["cnt"] == c["golden_exceptions"] and bw == c["golden_bw"], c["name"]
        assert (r["e"], r["f"], int(base)) == (c["e"], c["f"], c["base"]), c["name"]
        assert _same(r["enc"], g[cid + "_enc"]) and _same(r["exc"], g[cid + "_exc"]) and _same(r["pos"], g[cid + "_pos"]), c["name"]
        packed = port.ffor(r["enc"].view(_ut(x)), bw, int(base))
        assert _same(packed, g[cid + "_packed"]), c["name"]
        dec = port.patch(port.falp(packed, bw, int(base), r["f"], r["e"], x.dtype.itemsize), r["exc"], r["pos"])
        assert _same(dec, x), c["name"]


def test_port_reproduces_every_golden_rd_case(golden_vectors, port):
    g = golden_vectors
    cases = g.cases(scheme=1)
    assert len(cases) == 5
    for c in cases:
        cid, x, ref_state = c["id"], g[c["id"] + "_input"], g[c["id"] + "_state"]
        st = port.init(x)
        for key in ("scheme", "right_bw", "left_bw", "dict_size"):
            assert _same(st[key], ref_state[key]), (c["name"], key)
        # with the reference's state the streams are identical
        r = port.rd_encode(x, ref_state)
        assert _same(r["right"], g[cid + "_right"]) and _same(r["left"], g[cid + "_left"]), c["name"]
        assert _same(r["exc"], g[cid + "_exc"]) and _same(r["pos"], g[cid + "_pos"]), c["name"]
        pr, pl = port.ffor(r["right"], c["right_bw"], 0), port.ffor(r["left"], c["left_bw"], 0)
        assert _same(pr, g[cid + "_packed_right"]) and _same(pl, g[cid + "_packed_left"]), c["name"]
        dec = port.rd_decode(port.unffor(pr, c["right_bw"], 0, _ut(x)), port.unffor(pl, c["left_bw"], 0, np.uint16), r["exc"], r["pos"], ref_state)
        assert _same(dec, x), c["name"]
        # with its own state (own tie rule) it is still lossless
        r2 = port.rd_encode(x, st)
        assert _same(port.rd_decode(r2["right"], r2["left"], r2["exc"], r2["pos"], st), x), c["name"]


@pytest.mark.parametrize("name", ["city_temperature_f_tw", "food_prices_tw", "gov26_tw"])
def test_port_reproduces_golden_columns(name, golden_columns, port):
    gc = golden_columns
    x = gc[name + "_input"]
    want = host_column_from_golden(gc, name)
    got = port.encode_column(x, n_threads=4)
    assert got.meta.tobytes() == want.meta.tobytes()
    assert got.packed[: got.packed_bytes].tobytes() == want.packed[: want.packed_bytes].tobytes()
    assert got.exc_val[: got.n_exceptions].tobytes() == want.exc_val[: want.n_exceptions].tobytes()
    assert got.exc_pos[: got.n_exceptions].tobytes() == want.exc_pos[: want.n_exceptions].tobytes()
    assert port.decode_column(want, n_threads=4).tobytes() == x.tobytes()
    states = np.concatenate([port.init(x, off) for off in range(0, x.shape[0], 102400)])
    assert states.tobytes() == gc[name + "_states"].tobytes()
    entry = [e for e in gc.index if e["name"] == name][0]
    assert abs(8.0 * got.compressed_bytes() / x.shape[0] - entry["bits_per_value"]) < 1e-9


@pytest.mark.parametrize("kind,name", [(2, "synthetic_decimal_f64"), (3, "synthetic_highprec_f64"), (4, "synthetic_mixed_f32")])
def test_port_reproduces_synthetic_columns(kind, name, golden_columns, port):
    from oracle import pyoracle

    gc = golden_columns
    entry = [e for e in gc.index if e["name"] == name][0]
    x = pyoracle.generate(entry["n_values"], kind)
    assert [float(v) for v in x[:4]] == entry["first_values"]
    assert int(np.bitwise_xor.reduce(x.view(np.uint64 if kind != 4 else np.uint32))) == entry["xor_checksum"]
    got = port.encode_column(x, n_threads=4)
    want_meta = gc[name + "_meta"]
    assert got.packed_bytes == int(gc[name + "_totals"][0])
    # (for ALP_RD the number of exceptions depends on which of several equally frequent left parts made it into the
    # dictionary — STL-defined in the reference, rd.hpp:35-54 — so it is only compared for ALP columns)
    assert kind == 3 or got.n_exceptions == int(gc[name + "_totals"][1])
    if kind != 3:
        assert got.meta.tobytes() == want_meta.tobytes()
        sha = entry["sha256"]
        assert hashlib.sha256(got.packed[: got.packed_bytes].tobytes()).hexdigest() == sha["packed"]
        assert hashlib.sha256(got.exc_val[: got.n_exceptions].tobytes()).hexdigest() == sha["exc_val"]
        assert hashlib.sha256(got.exc_pos[: got.n_exceptions].tobytes()).hexdigest() == sha["exc_pos"]
    else:
        for key in ("scheme", "bw", "e", "f", "packed_off"):
            assert np.array_equal(got.meta[key], want_meta[key]), key
    assert port.decode_column(got, n_threads=4).tobytes() == x.tobytes()
    assert abs(8.0 * got.compressed_bytes() / x.shape[0] - entry["bits_per_value"]) < (1e-9 if kind != 3 else 0.05)


# ---------------------------------------------------------------------------------------------------------------------
# 2. restatement vs the compiled reference
# ---------------------------------------------------------------------------------------------------------------------
def test_float_fact_out_of_bounds(reference, port):
    """Constants<float>::FACT_ARR[10] is read out of bounds by the reference (decoder.hpp:129); the restatement's
    entry must be what the compiled reference sees."""
    for enc in (1, 7, -3):
        assert reference.decode_value(enc, 10, 0, 4) == port.decode_value(enc, 10, 0, 4)


def test_safe_sentinel(reference, port):
    """encode_value<true> returns (ST)ENCODING_UPPER_LIMIT for impossible values (encoder.hpp:84-86)."""
    for vb in (8, 4):
        for v in (float("nan"), float("inf"), -0.0, 1e300 if vb == 8 else 3e38):
            assert reference.encode_value(v, 0, 0, vb) == port.encode_value(v, 0, 0, vb), (vb, v)


def test_value_round_trip_fuzz(reference, port):
    rng = np.random.default_rng(0)
    for vb, max_e in ((8, 18), (4, 10)):
        dt = np.float64 if vb == 8 else np.float32
        vals = np.concatenate(
            [
                (rng.integers(-(10**9), 10**9, 3000) / 10.0 ** rng.integers(0, 8, 3000)).astype(dt),
                rng.integers(0, 1 << (8 * vb - 1), 1000, dtype=np.uint64).astype(np.uint64 if vb == 8 else np.uint32).view(dt),
                np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 9.3e18, -9.3e18, 2147483648.0, 4294967296.0, 0.1, 1e-300 if vb == 8 else 1e-40], dtype=dt),
            ]
        )
        for v in vals:
            e = int(rng.integers(0, max_e + 1))
            f = int(rng.integers(0, e + 1))
            a, b = reference.encode_value(float(v), f, e, vb), port.encode_value(float(v), f, e, vb)
            assert a == b, (vb, v, e, f)
            da, db = reference.decode_value(a, f, e, vb), port.decode_value(b, f, e, vb)
            assert np.array([da], dtype=dt).tobytes() == np.array([db], dtype=dt).tobytes(), (vb, v, e, f)


@pytest.mark.parametrize("tbits", [64, 32, 16, 8])
def test_ffor_unffor_every_width_vs_reference(tbits, reference, port):
    rng = np.random.default_rng(tbits)
    dt = {64: np.uint64, 32: np.uint32, 16: np.uint16, 8: np.uint8}[tbits]
    for bw in range(0, tbits + 1):
        base = int(rng.integers(0, 1 << min(tbits, 62)))
        vals = rng.integers(0, 1 << min(tbits, 63), size=1024, dtype=np.uint64).astype(dt)  # ffor masks, whatever the input
        a, b = reference.ffor(vals, bw, base), port.ffor(vals, bw, base)
        assert _same(a, b), (tbits, bw)
        assert _same(reference.unffor(a, bw, base, dt), port.unffor(b, bw, base, dt)), (tbits, bw)


def test_vector_encode_fuzz_vs_reference(reference, port):
    """Random decimal-ish, mixed and adversarial vectors through init → encode → analyze → ffor → falp → patch."""
    rng = np.random.default_rng(5)
    for trial in range(60):
        vb = 8 if trial % 3 else 4
        dt = np.float64 if vb == 8 else np.float32
        decimals = int(rng.integers(0, 7))
        x = (rng.integers(-(10 ** int(rng.integers(1, 9))), 10 ** int(rng.integers(1, 9)), 1024) / 10.0**decimals).astype(dt)
        n_bad = int(rng.integers(0, 200)) if trial % 2 else 0
        idx = rng.choice(1024, n_bad, replace=False)
        x[idx] = rng.integers(0, 1 << (8 * vb - 1), n_bad, dtype=np.uint64).astype(np.uint64 if vb == 8 else np.uint32).view(dt)
        if trial % 7 == 0:
            x[:3] = [np.nan, -0.0, np.inf]
        sa, sb = reference.init(x), port.init(x)
        for key in ("scheme", "k", "combos", "right_bw", "left_bw", "dict_size"):
            assert _same(sa[key], sb[key]), (trial, key)
        if int(sa["scheme"][0]) == 2:
            a, b = reference.encode(x, sa), port.encode(x, sa)
            for key in a:
                assert _same(a[key], b[key]), (trial, key)
            assert tuple(int(v) for v in reference.analyze_ffor(a["enc"])) == tuple(int(v) for v in port.analyze_ffor(b["enc"]))
        else:
            a, b = reference.rd_encode(x, sa), port.rd_encode(x, sa)
            for key in a:
                assert _same(a[key], b[key]), (trial, key)
            assert _same(reference.rd_decode(a["right"], a["left"], a["exc"], a["pos"], sa), x)
            assert _same(port.rd_decode(b["right"], b["left"], b["exc"], b["pos"], sa), x)


def test_falp_full_width_quirk_is_not_copied(reference, port):
    """The reference's fused kernel at bw == lane width multiplies the unsigned integer by frac10
    (src/falp.cpp:33311-33319); unffor + decode is the correct path there and is what the oracle restates."""
    rng = np.random.default_rng(9)
    enc = rng.integers(-(1 << 62), 1 << 62, 1024, dtype=np.int64)
    packed = reference.ffor(enc.view(np.uint64), 64, 0)
    unfused = reference.decode(reference.unffor(packed, 64, 0, np.uint64).view(np.int64), 2, 5)
    assert _same(port.falp(packed, 64, 0, 2, 5), unfused)


def test_column_drivers_agree(reference, port):
    from oracle import pyoracle

    for kind in (2, 4):
        x = pyoracle.generate(102400 + 13 * 1024, kind)
        a, b = reference.encode_column(x, n_threads=4), port.encode_column(x, n_threads=4)
        assert a.meta.tobytes() == b.meta.tobytes()
        assert a.packed[: a.packed_bytes].tobytes() == b.packed[: b.packed_bytes].tobytes()
        assert a.exc_val[: a.n_exceptions].tobytes() == b.exc_val[: b.n_exceptions].tobytes()
        assert reference.decode_column(b, n_threads=2).tobytes() == x.tobytes()
        assert port.decode_column(a, n_threads=2).tobytes() == x.tobytes()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_port_matches_reference_at_the_integer_overflow_boundaries(dtype, reference, port):
    """The column the GPU encoder's fast path is checked on (tests/test_gpu_columns.py): the C restatement and the
    compiled reference must agree on it byte for byte — x86 cast of out-of-range values, wrapping integer products, the
    SAFE sentinel in the sampling — so that either can be the judge."""
    from conftest import overflow_boundary_column

    x = overflow_boundary_column(dtype, np.random.default_rng(5))
    a, b = port.encode_column(x, n_threads=4), reference.encode_column(x, n_threads=4)
    alp = b.meta["scheme"] == 2
    assert alp.any() and np.array_equal(a.meta["scheme"], b.meta["scheme"])
    for key in ("exc_cnt", "bw", "e", "f"):
        assert np.array_equal(a.meta[key][alp], b.meta[key][alp]), key
    assert np.array_equal(a.meta["base"][alp], b.meta["base"][alp])
    if alp.all():
        assert a.packed[: a.packed_bytes].tobytes() == b.packed[: b.packed_bytes].tobytes()
        assert a.exc_val[: a.n_exceptions].tobytes() == b.exc_val[: b.n_exceptions].tobytes()
        assert a.exc_pos[: a.n_exceptions].tobytes() == b.exc_pos[: b.n_exceptions].tobytes()
    assert port.decode_column(b, n_threads=4).tobytes() == x.tobytes() and reference.decode_column(a, n_threads=4).tobytes() == x.tobytes()
