"""GPU parity, one vector at a time, through the C ABI's alpb200_prim_* entry points (the reference's primitive API).

Mirrors the reference's own test (test/test_alp_sample.cpp:97-187): for the first 1024 values of each fixture column —
init, scheme switch, encode → analyze_ffor → ffor, then falp + patch_exceptions (or the ALP_RD path), with the two
golden asserts of that test (bit_width, exceptions_count; :178-179) plus byte equality of every intermediate with
what the unmodified reference produced (tests/golden/reference_vectors.npz).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same(a, b):
    return np.asarray(a).tobytes() == np.asarray(b).tobytes()


def _ut(x):
    return np.uint64 if x.dtype.itemsize == 8 else np.uint32


def test_alp_cases_match_reference(golden_vectors):
    from alp_b200 import primitives as gpu

    g = golden_vectors
    cases = g.cases(scheme=2)
    assert len(cases) == 98
    for c in cases:
        cid, x, st = c["id"], g[c["id"] + "_input"], g[c["id"] + "_state"]
        r = gpu.encode(x, st)
        assert (r["e"], r["f"], r["cnt"]) == (c["e"], c["f"], c["cnt"]), c["name"]
        assert r["cnt"] == c["golden_exceptions"], c["name"]  # test_alp_sample.cpp:178
        assert _same(r["enc"], g[cid + "_enc"]), c["name"]
        assert _same(r["exc"], g[cid + "_exc"]) and _same(r["pos"], g[cid + "_pos"]), c["name"]
        bw, base = gpu.analyze_ffor(r["enc"])
        assert bw == c["golden_bw"] == c["bw"] and int(base) == c["base"], c["name"]  # test_alp_sample.cpp:179
        packed = gpu.ffor(r["enc"].view(_ut(x)), bw, int(base))
        assert _same(packed, g[cid + "_packed"]), c["name"]
        # fused decode + patch (test_alp_sample.cpp:169-170), and the unfused path (benchmark.cpp:129-131)
        dec = gpu.patch(gpu.falp(packed, bw, int(base), r["f"], r["e"], x.dtype.itemsize), r["exc"], r["pos"])
        assert _same(dec, x), c["name"]
        unpacked = gpu.unffor(packed, bw, int(base), _ut(x))
        assert _same(unpacked, r["enc"].view(_ut(x))), c["name"]
        dec2 = gpu.patch(gpu.decode(unpacked.view(r["enc"].dtype), r["f"], r["e"]), r["exc"], r["pos"])
        assert _same(dec2, x), c["name"]


def test_rd_cases_match_reference(golden_vectors):
    from alp_b200 import primitives as gpu

    g = golden_vectors
    cases = g.cases(scheme=1)
    assert len(cases) == 5
    for c in cases:
        cid, x, st = c["id"], g[c["id"] + "_input"], g[c["id"] + "_state"]
        r = gpu.rd_encode(x, st)
        assert r["cnt"] == c["cnt"], c["name"]
        assert _same(r["right"], g[cid + "_right"]) and _same(r["left"], g[cid + "_left"]), c["name"]
        assert _same(r["exc"], g[cid + "_exc"]) and _same(r["pos"], g[cid + "_pos"]), c["name"]
        pr = gpu.ffor(r["right"], c["right_bw"], 0)
        pl = gpu.ffor(r["left"], c["left_bw"], 0)
        assert _same(pr, g[cid + "_packed_right"]) and _same(pl, g[cid + "_packed_left"]), c["name"]
        dec = gpu.rd_decode(gpu.unffor(pr, c["right_bw"], 0, _ut(x)), gpu.unffor(pl, c["left_bw"], 0, np.uint16), r["exc"], r["pos"], st)
        assert _same(dec, x), c["name"]


def test_init_matches_reference_state(golden_vectors, port):
    """alpb200_prim_init (device sampling + (e,f) search + scheme decision + RD dictionary) against the reference's
    init on every fixture vector.  Scheme, top-k list, cut position, widths and dictionary size must equal the
    reference's; the dictionary *contents* among equally frequent left parts are STL-defined in the reference
    (rd.hpp:35-54), so those are compared with the restatement, which shares this library's tie rule."""
    from alp_b200 import primitives as gpu

    g = golden_vectors
    for c in g.index:
        x, ref = g[c["id"] + "_input"], g[c["id"] + "_state"]
        st = gpu.init(x)
        for key in ("scheme", "k", "combos", "right_bw", "left_bw", "dict_size"):
            assert _same(st[key], ref[key]), (c["name"], key, st[key], ref[key])
        assert _same(st, port.init(x)), c["name"]
        if c["scheme"] == 1:
            r = gpu.rd_encode(x, st)
            dec = gpu.rd_decode(r["right"], r["left"], r["exc"], r["pos"], st)
            assert _same(dec, x), c["name"]


@pytest.mark.parametrize("tbits", [64, 32, 16, 8])
def test_ffor_unffor_every_width(tbits, checker):
    """Every bit width of every lane width against the checker (src/fastlanes_generated_{ffor,unffor}.cpp dispatch)."""
    from alp_b200 import primitives as gpu

    rng = np.random.default_rng(tbits)
    dt = {64: np.uint64, 32: np.uint32, 16: np.uint16, 8: np.uint8}[tbits]
    for bw in range(0, tbits + 1):
        base = int(rng.integers(0, 1 << min(tbits, 62)))
        span = (1 << bw) - 1
        vals = (rng.integers(0, span + 1, size=1024, dtype=np.uint64) if bw < 64 else rng.integers(0, 1 << 63, size=1024, dtype=np.uint64) * 2 + 1)
        vals = ((vals + base) & ((1 << tbits) - 1)).astype(dt)
        want = checker.ffor(vals, bw, base)
        got = gpu.ffor(vals, bw, base)
        assert _same(got, want), (tbits, bw)
        back = gpu.unffor(got, bw, base, dt)
        assert _same(back, checker.unffor(want, bw, base, dt)), (tbits, bw)
        if bw > 0:
            assert _same(back, vals), (tbits, bw)


def test_ffor_rejects_oversized_width():
    import alp_b200
    from alp_b200 import primitives as gpu

    with pytest.raises(alp_b200.AlpError):
        gpu.ffor(np.zeros(1024, dtype=np.uint32), 33, 0)


def test_falp_every_width_random(checker):
    """Fused unpack+decode for every width and a spread of (e,f) against unffor + decoder::decode of the checker."""
    from alp_b200 import primitives as gpu

    rng = np.random.default_rng(7)
    for vb, ut, it, max_e in ((8, np.uint64, np.int64, 18), (4, np.uint32, np.int32, 10)):
        t = vb * 8
        for bw in range(0, t + 1):
            e = int(rng.integers(0, max_e + 1))
            f = int(rng.integers(0, e + 1))
            if vb == 4 and f == 10:
                f = 9
            base = int(rng.integers(-(1 << 40), 1 << 40)) if vb == 8 else int(rng.integers(-(1 << 20), 1 << 20))
            vals = rng.integers(0, 1 << min(bw, 62), size=1024, dtype=np.uint64) if bw else np.zeros(1024, dtype=np.uint64)
            enc = ((vals.astype(object) + base) % (1 << t)).astype(ut) if vb == 8 else ((vals + (base % (1 << 32))) & 0xFFFFFFFF).astype(ut)
            packed = checker.ffor(enc, bw, base % (1 << t))
            want = checker.decode(checker.unffor(packed, bw, base % (1 << t), ut).view(it), f, e)
            got = gpu.falp(packed, bw, base % (1 << t), f, e, vb)
            assert _same(got, want), (vb, bw, e, f)


def test_special_values_and_casts(checker):
    """NaN / ±Inf / -0.0 / huge magnitudes: exception selection must follow the reference's x86 cast semantics."""
    from alp_b200 import primitives as gpu

    rng = np.random.default_rng(11)
    for dt in (np.float64, np.float32):
        x = (rng.integers(0, 100000 if dt == np.float64 else 5000, size=1024) / 100.0).astype(dt)
        specials = [np.nan, np.inf, -np.inf, -0.0, 0.0, 9.3e18, -9.3e18, 1e300 if dt == np.float64 else 1e38, 2147483648.0, -2147483649.0, 5e-324 if dt == np.float64 else 1e-45]
        for i, s in enumerate(specials):
            x[7 + 31 * i] = dt(s)
        st = checker.init(x)
        assert int(st["scheme"][0]) == 2
        a, b = gpu.encode(x, st), checker.encode(x, st)
        for key in a:
            assert _same(a[key], b[key]), (dt, key)
        bw, base = gpu.analyze_ffor(a["enc"])
        assert (bw, int(base)) == tuple(int(v) for v in checker.analyze_ffor(b["enc"]))
        ut = np.uint64 if dt == np.float64 else np.uint32
        packed = gpu.ffor(a["enc"].view(ut), bw, int(base))
        dec = gpu.patch(gpu.falp(packed, bw, int(base), a["f"], a["e"], x.dtype.itemsize), a["exc"], a["pos"])
        assert _same(dec, x)


def test_all_exceptions_and_leading_exceptions(checker):
    """Fill value = first non-exception's encoded integer, 0 when there is none (encoder.hpp:382-388)."""
    from alp_b200 import primitives as gpu

    x = np.full(1024, np.nan)
    st = np.zeros(1, dtype=checker.init(np.arange(1024.0)).dtype)
    st["scheme"], st["k"] = 2, 1
    st["combos"][0, 0] = (14, 12)
    for variant in range(3):
        y = x.copy()
        if variant == 1:
            y[1000:] = 12.5
        if variant == 2:
            y[0:5] = np.inf
            y[5:] = np.arange(1019) / 4.0
        a, b = gpu.encode(y, st), checker.encode(y, st)
        for key in a:
            assert _same(a[key], b[key]), (variant, key)
