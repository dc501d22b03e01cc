"""GPU parity of the batched hot path (alpb200_rowgroup_init / encode / decode) on whole columns.

Ground truth is (a) the committed golden fixtures produced by the unmodified reference (tests/golden/) and (b) the CPU
checker that travels with the repository (oracle/_ref when built, else the C restatement).  Integer / byte work must
be bit exact: metadata records, packed blocks, exception values and positions, and the decoded values.
"""
import hashlib

import numpy as np
import pytest

from conftest import host_column_from_golden

pytestmark = pytest.mark.gpu


def _dev():
    import torch

    return torch.device("cuda:0")


def _bits(t):
    import torch

    return t.view(torch.int64 if t.element_size() == 8 else torch.int32)


def _meta_fields_equal(a, b, name):
    for key in ("packed_off", "exc_off", "exc_cnt", "scheme", "bw", "e", "f"):
        assert np.array_equal(a[key], b[key]), (name, key)
    alp = a["scheme"] == 2
    assert np.array_equal(a["base"][alp], b["base"][alp]), (name, "base")
    assert a[~alp].tobytes() == b[~alp].tobytes() or np.array_equal(
        a[~alp].view(np.uint8).reshape(-1, 32)[:, :16], b[~alp].view(np.uint8).reshape(-1, 32)[:, :16]
    ), (name, "rd_dict")


def _assert_columns_equal(got, want, name):
    """got, want: HostColumn.  Everything the decoder reads must be byte-identical."""
    assert got.n_vectors == want.n_vectors
    assert got.packed_bytes == want.packed_bytes and got.n_exceptions == want.n_exceptions, name
    _meta_fields_equal(got.meta, want.meta, name)
    assert got.packed[: got.packed_bytes].tobytes() == want.packed[: want.packed_bytes].tobytes(), (name, "packed")
    n = got.n_exceptions
    assert got.exc_pos[:n].tobytes() == want.exc_pos[:n].tobytes(), (name, "exc_pos")
    assert got.exc_val[:n].tobytes() == want.exc_val[:n].tobytes(), (name, "exc_val")


@pytest.mark.parametrize("name", ["city_temperature_f_tw", "food_prices_tw", "gov26_tw"])
def test_real_columns_against_golden(name, golden_columns):
    """128 vectors = 2 row-groups of real data with many distinct (bw, e, f) per column."""
    import torch

    import alp_b200

    gc = golden_columns
    x = gc[name + "_input"]
    ref_col = host_column_from_golden(gc, name)
    xd = torch.from_numpy(x).to(_dev())

    # decode of the reference-encoded column
    y = alp_b200.decode(alp_b200.DeviceColumn.from_host(ref_col, _dev()))
    assert y.cpu().numpy().tobytes() == x.tobytes()

    # device init == reference init (ALP columns: no STL-dependent ties)
    states = alp_b200.rowgroup_init(xd).cpu().numpy().view(alp_b200._abi.RG_STATE_DTYPE).reshape(-1)
    assert states.tobytes() == gc[name + "_states"].tobytes()

    # device encode == reference encode, byte for byte
    col = alp_b200.encode(xd)
    _assert_columns_equal(col.to_host(), ref_col, name)
    assert torch.equal(_bits(alp_b200.decode(col)), _bits(xd))

    # partial decode of a vector range
    part = alp_b200.decode(col, first=37, n=50)
    assert torch.equal(_bits(part), _bits(xd[37 * 1024 : 87 * 1024]))


@pytest.mark.parametrize("kind,name", [(2, "synthetic_decimal_f64"), (3, "synthetic_highprec_f64"), (4, "synthetic_mixed_f32")])
def test_synthetic_columns_against_golden(kind, name, golden_columns, port):
    """The three synthetic generators of SURVEY.md §8d, 4 row-groups each: generator, init, metadata and payload
    digests against what the reference produced."""
    import torch

    import alp_b200
    from oracle import pyoracle

    gc = golden_columns
    entry = [e for e in gc.index if e["name"] == name][0]
    n = entry["n_values"]
    xd = alp_b200.generate(n, kind, _dev())
    x = xd.cpu().numpy()
    assert x.tobytes() == pyoracle.generate(n, kind).tobytes()
    assert [float(v) for v in x[:4]] == entry["first_values"]
    assert int(np.bitwise_xor.reduce(x.view(np.uint64 if kind != 4 else np.uint32))) == entry["xor_checksum"]

    states = alp_b200.rowgroup_init(xd)
    col = alp_b200.encode(xd, states)
    h = col.to_host()
    want_meta = gc[name + "_meta"]
    assert h.packed_bytes == int(gc[name + "_totals"][0])
    # (for ALP_RD the number of exceptions depends on which of several equally frequent left parts made it into the
    # dictionary — STL-defined in the reference, rd.hpp:35-54 — so it is only compared for ALP columns)
    assert kind == 3 or h.n_exceptions == int(gc[name + "_totals"][1])
    if kind != 3:
        # ALP: everything is determined — metadata and payload digests equal the reference's
        _meta_fields_equal(h.meta, want_meta, name)
        sha = entry["sha256"]
        assert hashlib.sha256(h.packed[: h.packed_bytes].tobytes()).hexdigest() == sha["packed"]
        assert hashlib.sha256(h.exc_val[: h.n_exceptions].tobytes()).hexdigest() == sha["exc_val"]
        assert hashlib.sha256(h.exc_pos[: h.n_exceptions].tobytes()).hexdigest() == sha["exc_pos"]
        st = states.cpu().numpy().view(alp_b200._abi.RG_STATE_DTYPE).reshape(-1)
        assert st.tobytes() == gc[name + "_states"].tobytes()
    else:
        # ALP_RD: cut, widths, dictionary size and sizes equal the reference's; dictionary order on ties is this
        # library's rule (shared with the restatement), so compare the full stream with the restatement
        for key in ("scheme", "bw", "e", "f", "packed_off"):
            assert np.array_equal(h.meta[key], want_meta[key]), key
        _assert_columns_equal(h, port.encode_column(x), name)
    assert torch.equal(_bits(alp_b200.decode(col)), _bits(xd))


def test_cross_decode_with_checker(checker):
    """GPU-encoded columns decode on the CPU checker and vice versa (mixed ALP / ALP_RD row-groups in one column)."""
    import torch

    import alp_b200
    from oracle import pyoracle

    x = np.concatenate([pyoracle.generate(102400, 2), pyoracle.generate(102400, 3), pyoracle.generate(51200, 2, first_index=7)])
    xd = torch.from_numpy(x).to(_dev())
    col = alp_b200.encode(xd)
    h = col.to_host()
    assert set(h.meta["scheme"].tolist()) == {1, 2}
    assert checker.decode_column(h).tobytes() == x.tobytes()
    ref_col = checker.encode_column(x)
    y = alp_b200.decode(alp_b200.DeviceColumn.from_host(ref_col, _dev()))
    assert y.cpu().numpy().tobytes() == x.tobytes()
    assert torch.equal(_bits(alp_b200.decode(col)), _bits(xd))


def test_reference_state_gives_reference_stream_for_rd(reference):
    """Fed the reference's own row-group states (dictionary order and exception indices included), the device encoder
    reproduces the reference's ALP_RD stream byte for byte."""
    import torch

    import alp_b200
    from oracle import pyoracle

    x = pyoracle.generate(3 * 102400, 3)
    states = np.concatenate([reference.init(x, off) for off in range(0, x.shape[0], 102400)])
    sd = torch.from_numpy(states.view(np.uint8).reshape(len(states), -1)).to(_dev())
    col = alp_b200.encode(torch.from_numpy(x).to(_dev()), sd)
    _assert_columns_equal(col.to_host(), reference.encode_column(x), "rd-with-reference-state")


def test_edge_shapes(checker):
    """One vector, a ragged last row-group, all-equal values, all exceptions, every special value."""
    import torch

    import alp_b200

    rng = np.random.default_rng(3)
    cols = {
        "one_vector": np.round(rng.normal(20, 5, 1024), 2),
        "ragged_rowgroup": np.round(rng.normal(20, 5, 1024 * 137), 1),
        "constant": np.full(1024 * 3, 10.23),
        "all_nan": np.full(1024 * 2, np.nan),
        "zeros_f32": np.zeros(1024 * 2, dtype=np.float32),
        "specials_f32": np.tile(np.array([np.nan, np.inf, -np.inf, -0.0, 0.0, 1.5, 3e38, 2147483648.0], dtype=np.float32), 128 * 3),
        "random_bits_f64": rng.integers(0, 1 << 63, size=1024 * 101, dtype=np.uint64).view(np.float64),
        "random_bits_f32": rng.integers(0, 1 << 32, size=1024 * 101, dtype=np.uint64).astype(np.uint32).view(np.float32),
    }
    for name, x in cols.items():
        xd = torch.from_numpy(x).to(_dev())
        col = alp_b200.encode(xd)
        assert torch.equal(_bits(alp_b200.decode(col)), _bits(xd)), name
        h = col.to_host()
        assert checker.decode_column(h).tobytes() == x.tobytes(), name
        want = checker.encode_column(x)
        if set(want.meta["scheme"].tolist()) == {2}:
            _assert_columns_equal(h, want, name)


def _width_sweep_column(dtype, widths, rng, n_exceptions=3):
    """One ROW-GROUP (100 vectors) per bit width w, so that the row-group's (e,f) search sees only that kind of data:
    integers spanning exactly [0, 2^w - 1] in every vector (FFOR then packs w bits), plus a few NaNs per vector (exceptions)."""
    mant = 52 if dtype == np.float64 else 23
    parts = []
    for w in widths:
        span = (1 << w) - 1
        step = 1 if w <= mant else 1 << (w - mant)  # keep every value exactly representable
        k = rng.integers(0, span // step + 1, size=(100, 1024), dtype=np.uint64)
        k[:, 0], k[:, 1] = 0, span // step
        v = (k * np.uint64(step)).astype(dtype)
        v = rng.permuted(v, axis=1)
        for row in v:
            row[rng.choice(1024, size=n_exceptions, replace=False)] = np.nan
        parts.append(v.reshape(-1))
    return np.concatenate(parts)


# (wider integer columns fall to ALP_RD: the row-group search gives up at 48 / 22 bits, constants.hpp:33,69)
@pytest.mark.parametrize("dtype,widths", [(np.float64, list(range(0, 48))), (np.float32, list(range(0, 22)))])
def test_batched_encode_every_bit_width(dtype, widths, checker):
    """Every FFOR width through the BATCHED kernels (in-place packing + bulk store for narrow blocks, direct line
    stores for wide ones; the width-specialised unpackers; the integer fast path of the fused SUM): byte-identical to
    the checker's column, decodable both ways, and per-vector SUM within 1e-12 of the decoded values' sum."""
    import torch

    import alp_b200

    rng = np.random.default_rng(11)
    x = _width_sweep_column(dtype, widths, rng)
    xd = torch.from_numpy(x).to(_dev())
    col = alp_b200.encode(xd)
    h = col.to_host()
    want = checker.encode_column(x, n_threads=8)
    assert set(want.meta["scheme"].tolist()) == {2}
    assert sorted(set(want.meta["bw"].tolist())) == widths  # the sweep really covers every width
    _assert_columns_equal(h, want, "width-sweep")
    y = alp_b200.decode(col)
    assert torch.equal(_bits(y), _bits(xd))
    assert checker.decode_column(h, n_threads=8).tobytes() == x.tobytes()
    # SUM of one vector per width (NaN exceptions propagate, so compare on a NaN-free copy of the column)
    clean = np.nan_to_num(x, nan=1.0)
    cd = torch.from_numpy(clean).to(_dev())
    ccol = alp_b200.encode(cd)
    ccol.read_totals()
    for i, w in enumerate(widths):
        first = i * 100 + 7
        v = cd[first * 1024 : (first + 1) * 1024].double()
        got = float(alp_b200.decode_sum(ccol, first=first, n=1).item())
        assert abs(got - float(v.sum().item())) <= 1e-12 * max(float(v.abs().sum().item()), 1e-300), (w, got, float(v.sum().item()))
    total = float(alp_b200.decode_sum(ccol).item())
    assert abs(total - float(cd.double().sum().item())) <= 1e-12 * float(cd.double().abs().sum().item())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_encode_near_the_integer_overflow_boundaries(dtype, checker):
    """The encoder's floating-point fast path (alp_encode.cuh, analyze_rows<FAST>) must hand over to the exact recipe
    wherever enc * 10^f can wrap or the x86 cast semantics matter: decimals whose scaled integer sits right at
    +-2^63 / 10^f (+-2^31 / 10^f for floats), +-2^63 itself, huge values, infinities and NaNs, mixed into ordinary
    decimal vectors.  Byte-identical to the checker's column, for several decimal scales."""
    import torch

    import alp_b200

    from conftest import overflow_boundary_column

    x = overflow_boundary_column(dtype, np.random.default_rng(5))
    xd = torch.from_numpy(x).to(_dev())
    col = alp_b200.encode(xd)
    h = col.to_host()
    want = checker.encode_column(x, n_threads=8)
    assert 2 in set(want.meta["scheme"].tolist())
    _assert_columns_equal(h, want, "overflow-boundaries")
    assert torch.equal(_bits(alp_b200.decode(col)), _bits(xd))


def _blocks(h):
    """per-vector (packed block bytes, exception values, exception positions) of a HostColumn, whatever its layout"""
    m = h.meta[: h.n_vectors]
    units = np.where(m["scheme"] == 1, m["bw"].astype(np.int64) + m["e"].astype(np.int64), m["bw"].astype(np.int64))
    out = []
    for v in range(h.n_vectors):
        p, e, c = int(m["packed_off"][v]) * 128, int(m["exc_off"][v]), int(m["exc_cnt"][v])
        out.append((h.packed[p : p + int(units[v]) * 128].tobytes(), h.exc_val[e : e + c].tobytes(), h.exc_pos[e : e + c].tobytes()))
    return out, units


@pytest.mark.parametrize("kind", [2, 3, 4])
def test_completion_order_layout(kind, checker):
    """alpb200_encode_unordered_*: the same per-vector blocks, exception runs and record fields as the vector-order
    encoder (hence as the checker), dense (blocks tile [0, total) without gaps or overlaps), decodable by the GPU, by the
    CPU checker, through the host-buffer API, and shardable."""
    import torch

    import alp_b200
    from alp_b200 import shard
    from oracle import pyoracle

    n = 7 * 102400 + 31 * 1024
    x = pyoracle.generate(n, kind)
    xd = torch.from_numpy(x).to(_dev())
    ordered = alp_b200.encode(xd).to_host()
    col = alp_b200.encode(xd, ordered=False)
    h = col.to_host()
    assert (h.packed_bytes, h.n_exceptions) == (ordered.packed_bytes, ordered.n_exceptions)
    for key in ("exc_cnt", "scheme", "bw", "e", "f"):
        assert np.array_equal(h.meta[key], ordered.meta[key]), key
    got, units = _blocks(h)
    want, _ = _blocks(ordered)
    assert got == want
    # dense: sorted by offset, every non-empty block / exception run starts where the previous one ends
    def tiles(offsets, sizes, total):
        keep = sizes > 0
        off, sz = offsets[keep].astype(np.int64), sizes[keep].astype(np.int64)
        order = np.argsort(off)
        off, sz = off[order], sz[order]
        return off.size == 0 or (off[0] == 0 and np.array_equal(off[1:], (off + sz)[:-1]) and int(off[-1] + sz[-1]) == total)

    assert tiles(h.meta["packed_off"], units, h.packed_bytes // 128)
    assert tiles(h.meta["exc_off"], h.meta["exc_cnt"].astype(np.int64), h.n_exceptions)
    # decodable everywhere
    assert torch.equal(_bits(alp_b200.decode(col)), _bits(xd))
    assert checker.decode_column(h, n_threads=4).tobytes() == x.tobytes()
    got_sum = float(alp_b200.decode_sum(col).item())
    want_sum = float(xd.double().sum().item())
    assert abs(got_sum - want_sum) <= 1e-9 * max(1.0, abs(want_sum))
    # host-buffer API with the option set: compress -> decompress / sum
    codec = alp_b200.HostCodec(n // 1024, x.dtype.itemsize, ordered=False)
    hc = codec.compress(x)
    assert codec.decompress(hc).tobytes() == x.tobytes()
    assert abs(codec.sum(hc) - want_sum) <= 1e-9 * max(1.0, abs(want_sum))
    codec.close()
    # a shard cut out of it decodes on its own
    t = shard.column_tensors(col)
    s = shard.slice_column(t, 200, 300)
    sc = shard.tensors_to_device_column(s, x.dtype.itemsize, _dev())
    assert torch.equal(_bits(alp_b200.decode(sc)), _bits(xd[200 * 1024 : 500 * 1024]))


def test_host_decompress_of_a_shuffled_column(checker):
    """decompress_host / sum_host make no assumption about block order: a column whose blocks were permuted at random."""
    import alp_b200
    from conftest import shuffle_layout
    from oracle import pyoracle

    n = 40 * 102400
    x = pyoracle.generate(n, 2)
    shuffled = shuffle_layout(checker.encode_column(x, n_threads=8), np.random.default_rng(4))
    codec = alp_b200.HostCodec(n // 1024, 8)
    assert codec.decompress(shuffled).tobytes() == x.tobytes()
    assert abs(codec.sum(shuffled) - float(x.sum())) <= 1e-12 * float(np.abs(x).sum())
    codec.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_rowgroup_init_near_the_integer_overflow_boundaries(dtype, checker):
    """The (e,f) search of alpb200_rowgroup_init_* uses the same floating-point shortcut as the encoder; here the SAMPLED
    values themselves sit at the boundaries (products around +-2^63 | +-2^31, impossible values, specials), many per
    row-group.  Scheme, number of candidates and the candidate list must equal the checker's for every row-group."""
    import torch

    import alp_b200
    from alp_b200 import _abi

    rng = np.random.default_rng(21)
    top = 63 if dtype == np.float64 else 31
    n_rg = 48
    span = 3000 if dtype == np.float64 else 100  # (wider float ranges fall to ALP_RD, where there is nothing to search)
    x = np.round(rng.uniform(-span, span, size=(n_rg, 102400)), 2)
    for rg in range(n_rg):
        decimals, f = rg % 4, (rg * 5) % (14 if dtype == np.float64 else 8)
        edge = (2.0**top) / (10.0**f) / (10.0**decimals)
        cand = np.array([edge, -edge, np.nextafter(edge, 0), np.nextafter(edge, np.inf), np.floor(edge), -np.floor(edge) - 1, edge * (1 - 1e-7),
                         edge * (1 + 1e-7), edge / 2, edge * 2, 2.0**top, -(2.0**top), 1e30, -1e30, np.inf, np.nan, -0.0, 1e-320, -1e-320])
        hit = rng.random(102400) < (0.05 + 0.1 * (rg % 3))
        x[rg, hit] = rng.choice(cand, size=int(hit.sum()))
    with np.errstate(over="ignore"):
        x = x.reshape(-1).astype(dtype)
    n_alp = 0
    got = alp_b200.rowgroup_init(torch.from_numpy(x).to(_dev())).cpu().numpy().view(_abi.RG_STATE_DTYPE).reshape(-1)
    for rg in range(n_rg):
        want = checker.init(x, rg * 102400)[0]
        assert int(got[rg]["scheme"]) == int(want["scheme"]), rg
        if int(want["scheme"]) == 2:
            n_alp += 1
            k = int(want["k"])
            assert int(got[rg]["k"]) == k and np.array_equal(got[rg]["combos"][:k], want["combos"][:k]), (rg, got[rg]["combos"], want["combos"])
    assert n_alp >= n_rg // 2


@pytest.mark.parametrize("ordered", [True, False])
def test_rd_every_right_width_with_given_states(ordered, checker):
    """ALP_RD through the batched encoder for EVERY cut position (right widths 48..63) and index width 1..3, with
    hand-made row-group states: in-place packing of wide blocks (only the words that would land on unread rows are
    deferred), the direct path where block + index block outgrow the tile (62+3, 63+2, 63+3), left-part exceptions.
    The decoder (independent, width-specialised unpack) and the CPU checker must both give the values back, and the
    right-part blocks must equal the single-vector primitive's FFOR of the same right parts."""
    import torch

    import alp_b200
    from alp_b200 import _abi
    from alp_b200 import primitives as gpu

    rng = np.random.default_rng(17)
    combos = [(rbw, lbw) for rbw in range(48, 64) for lbw in (1, 2, 3)]
    states = np.zeros(len(combos), dtype=_abi.RG_STATE_DTYPE)
    parts = []
    for i, (rbw, lbw) in enumerate(combos):
        ds = 1 << lbw if lbw > 1 else 2
        dict_vals = rng.choice(1 << (64 - rbw), size=min(ds, 1 << (64 - rbw)), replace=False).astype(np.uint64)
        ds = dict_vals.size
        states[i]["scheme"], states[i]["right_bw"], states[i]["left_bw"], states[i]["dict_size"] = 1, rbw, lbw, ds
        states[i]["dict"][:ds] = dict_vals.astype(np.uint16)
        left = dict_vals[rng.integers(0, ds, size=102400)]
        stray = rng.random(102400) < 0.01  # left parts outside the dictionary: exceptions
        left[stray] = rng.integers(0, 1 << (64 - rbw), size=int(stray.sum()), dtype=np.uint64)
        right = rng.integers(0, 1 << 62, size=102400, dtype=np.uint64) & np.uint64((1 << rbw) - 1)
        parts.append((left << np.uint64(rbw)) | right)
    bits = np.concatenate(parts)
    x = bits.view(np.float64)
    xd = torch.from_numpy(x).to(_dev())
    sd = torch.from_numpy(states.view(np.uint8).reshape(len(states), -1)).to(_dev())
    col = alp_b200.encode(xd, sd, ordered=ordered)
    col.read_totals()
    assert torch.equal(_bits(alp_b200.decode(col)), _bits(xd))
    h = col.to_host()
    assert checker.decode_column(h, n_threads=8).tobytes() == x.tobytes()
    assert np.array_equal(h.meta["bw"][::100], np.array([c[0] for c in combos], dtype=np.uint8))
    for i, (rbw, lbw) in enumerate(combos):  # one vector per state against the primitive path (FFOR straight to memory)
        v = i * 100 + 3
        want = gpu.ffor(bits[v * 1024 : (v + 1) * 1024] & np.uint64((1 << rbw) - 1), rbw, 0)
        p0 = int(h.meta["packed_off"][v]) * 128
        assert h.packed[p0 : p0 + 128 * rbw].tobytes() == want.tobytes(), (rbw, lbw)


def test_capacity_overflow_is_reported():
    import torch

    import alp_b200

    x = alp_b200.generate(1024 * 64, 2, _dev())
    col = alp_b200.DeviceColumn(64, 8, _dev(), packed_capacity=1024, exc_capacity=4)
    alp_b200.encode(x, col=col)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError):
        col.read_totals()


def test_empty_and_bad_arguments():
    import torch

    import alp_b200

    with pytest.raises(ValueError):
        alp_b200.encode(torch.zeros(1000, dtype=torch.float64, device=_dev()))
    col = alp_b200.encode(alp_b200.generate(2048, 2, _dev()))
    with pytest.raises(alp_b200.AlpError):
        alp_b200.decode(col, first=1, n=5)
    assert alp_b200.decode(col, first=2, n=0).numel() == 0


@pytest.mark.parametrize("kind", [2, 3, 4])
def test_host_codec_round_trip(kind, checker):
    """alpb200_compress_host / decompress_host: host buffers in, host buffers out (copies inside the call)."""
    import alp_b200
    from oracle import pyoracle

    n = 3 * 102400 + 17 * 1024
    x = pyoracle.generate(n, kind)
    codec = alp_b200.HostCodec(n // 1024, x.dtype.itemsize)
    col = codec.compress(x)
    assert checker.decode_column(col).tobytes() == x.tobytes()
    assert codec.decompress(col).tobytes() == x.tobytes()
    assert codec.decompress(checker.encode_column(x)).tobytes() == x.tobytes()
    codec.close()


@pytest.mark.parametrize("kind,log2n", [(2, 30), (3, 30), (4, 28)])
def test_baseline_configs_at_full_size(kind, log2n, checker, port):
    """BASELINE.json configs 2-4 at their FULL sizes (2^30 f64 / 2^28 f32 values), through size-independent properties:
    encode -> decode is the identity (bit patterns, plus an XOR checksum of checksums), the metadata is consistent with
    the totals (dense offsets in vector order), both layouts hold the same blocks, the fused SUM agrees with the decoded
    column, and slices from the start, the middle and the ragged end of the column are byte-identical to what the CPU
    checker makes of the same values."""
    import torch

    import alp_b200
    from alp_b200 import _abi
    from oracle import pyoracle

    n = 1 << log2n
    xd = alp_b200.generate(n, kind, _dev())
    col = alp_b200.encode(xd)
    packed_bytes, n_exc = col.read_totals()
    y = alp_b200.decode(col)
    ib = torch.int64 if xd.element_size() == 8 else torch.int32
    assert torch.equal(y.view(ib), xd.view(ib))

    def xor_all(t):  # tree reduction of the bit patterns: a checksum of (chunk) checksums
        t = t.view(ib).clone()
        while t.numel() > 1:
            h = t.numel() // 2
            t = t[:h] ^ t[h : 2 * h] if t.numel() % 2 == 0 else torch.cat([t[:h] ^ t[h : 2 * h], t[2 * h :]])
        return int(t.item())

    assert xor_all(y) == xor_all(xd)
    want_sum = float(y.double().sum().item())
    del y
    meta = col.meta.cpu().numpy().view(_abi.VEC_META_DTYPE).reshape(-1)
    units = np.where(meta["scheme"] == 1, meta["bw"].astype(np.int64) + meta["e"], meta["bw"].astype(np.int64))
    cnt = meta["exc_cnt"].astype(np.int64)
    assert packed_bytes == int(units.sum()) * 128 and n_exc == int(cnt.sum())
    assert np.array_equal(meta["packed_off"].astype(np.int64), np.concatenate([[0], np.cumsum(units)[:-1]]))
    assert np.array_equal(meta["exc_off"].astype(np.int64), np.concatenate([[0], np.cumsum(cnt)[:-1]]))
    assert set(meta["scheme"].tolist()) == ({1} if kind == 3 else {2})
    got_sum = float(alp_b200.decode_sum(col).item())
    assert abs(got_sum - want_sum) <= 1e-9 * max(1.0, abs(want_sum))
    # slices against the CPU checker: first row-groups, a middle range, the short last row-group
    n_vec = n // 1024
    for first, count in ((0, 200), (n_vec // 2 // 100 * 100, 300), (n_vec // 100 * 100, n_vec % 100)):
        if count == 0:
            continue
        host = pyoracle.generate(count * 1024, kind, first_index=first * 1024)
        assert host.tobytes() == xd[first * 1024 : (first + count) * 1024].cpu().numpy().tobytes()
        judge = port if kind == 3 else checker
        _assert_columns_equal(col.to_host(first, count), judge.encode_column(host, n_threads=8), "full-size slice %d" % first)
    # the completion-order layout of the same column: same totals, same records apart from the offsets
    un = alp_b200.encode(xd, ordered=False)
    assert un.read_totals() == (packed_bytes, n_exc)
    um = un.meta.cpu().numpy().view(_abi.VEC_META_DTYPE).reshape(-1)
    for key in ("exc_cnt", "scheme", "bw", "e", "f"):
        assert np.array_equal(um[key], meta[key]), key
    z = alp_b200.decode(un)
    assert torch.equal(z.view(ib), xd.view(ib))


@pytest.mark.parametrize("kind,n_vec", [(2, 16 * 1000 + 3 * 100 + 41), (3, 2300), (4, 40 * 100)])
def test_pipelined_host_compress_equals_one_launch(kind, n_vec):
    """alpb200_compress_host pipelines the column in chunks of whole row-groups whose encodes APPEND to the column; the
    result must be byte for byte the column one device launch produces (several chunks, a ragged last row-group)."""
    import torch

    import alp_b200
    from oracle import pyoracle

    x = pyoracle.generate(n_vec * 1024, kind)
    codec = alp_b200.HostCodec(n_vec, x.dtype.itemsize)
    got = codec.compress(x)
    codec.close()
    want = alp_b200.encode(torch.from_numpy(x).to(_dev())).to_host()
    _assert_columns_equal(got, want, "pipelined-compress")
    assert int(got.totals[3]) == int(want.totals[3])


@pytest.mark.parametrize("kind,ordered", [(2, True), (3, True), (4, True), (2, False)])
def test_appending_encodes_equal_one_launch(kind, ordered):
    """alpb200_encode_ex_* with ALPB200_ENCODE_APPEND: a column encoded in three calls over consecutive row-group ranges.
    Vector-order layout: byte for byte the column of one call.  Completion order: the same blocks, and every call's
    blocks lie behind those of the previous calls."""
    import torch

    import alp_b200
    from oracle import pyoracle

    n_vec = 1000 + 700 + 345
    x = pyoracle.generate(n_vec * 1024, kind)
    xd = torch.from_numpy(x).to(_dev())
    want = alp_b200.encode(xd).to_host()
    col = alp_b200.DeviceColumn(n_vec, x.dtype.itemsize, _dev())
    ends = []
    for first, count in ((0, 1000), (1000, 700), (1700, 345)):
        alp_b200.encode(xd[first * 1024 : (first + count) * 1024], col=col, ordered=ordered, append_at=first)
        ends.append(col.read_totals())
    got = col.to_host()
    assert ends[-1] == (want.packed_bytes, want.n_exceptions)
    if ordered:
        _assert_columns_equal(got, want, "append")
    else:
        assert _blocks(got)[0] == _blocks(want)[0]
        off = got.meta["packed_off"].astype(np.int64) * 128
        assert off[:1000].max() < ends[0][0] <= off[1000:1700].min() and off[1000:1700].max() < ends[1][0] <= off[1700:].min()
    assert torch.equal(_bits(alp_b200.decode(col)), _bits(xd))


def test_large_column_properties():
    """2^26 values (BASELINE config 2 at 1/16 scale; the full size runs in bench.py, which verifies its round trip
    too): encode→decode is the identity, sizes match the per-vector metadata, positions are sorted, blocks are dense."""
    import torch

    import alp_b200

    n = 1 << 26
    xd = alp_b200.generate(n, 2, _dev())
    col = alp_b200.encode(xd)
    packed_bytes, n_exc = col.read_totals()
    y = alp_b200.decode(col)
    assert torch.equal(_bits(y), _bits(xd))
    meta = col.meta.cpu().numpy().view(alp_b200._abi.VEC_META_DTYPE).reshape(-1)
    units = meta["bw"].astype(np.int64)
    assert np.array_equal(meta["packed_off"].astype(np.int64), np.concatenate([[0], np.cumsum(units)[:-1]]))
    assert packed_bytes == int(units.sum()) * 128
    cnt = meta["exc_cnt"].astype(np.int64)
    assert np.array_equal(meta["exc_off"].astype(np.int64), np.concatenate([[0], np.cumsum(cnt)[:-1]]))
    assert n_exc == int(cnt.sum())
    assert set(meta["bw"].tolist()) == {20}  # every vector of the decimal column packs to 20 bits
    pos = col.exc_pos[:n_exc].cpu().numpy().view(np.uint16).astype(np.int64)
    owner = np.repeat(np.arange(meta.shape[0]), cnt)
    key = owner * 1024 + pos
    assert np.all(np.diff(key) > 0)  # ascending positions inside every vector, vectors in order


@pytest.mark.parametrize("kind", [2, 3, 4])
def test_fused_decode_sum(kind):
    """alpb200_decode_sum: decode + aggregate without materialising (reference: alp_func + aggr_plus, q1.cpp:63-102).
    Floating point: the order of additions differs from a sequential sum, so the check is relative (1e-9) against a
    float64 sum of the bit-exact decoded column."""
    import torch

    import alp_b200

    n = 5 * 102400 + 13 * 1024
    xd = alp_b200.generate(n, kind, _dev())
    col = alp_b200.encode(xd)
    col.read_totals()
    want = float(alp_b200.decode(col).double().sum().item())
    got = float(alp_b200.decode_sum(col).item())
    assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (kind, got, want)
    # and against a sum made on the checker's side, from the checker's own decode of the column
    from oracle import pyoracle

    decoded = pyoracle.best().decode_column(col.to_host()).astype(np.float64)
    assert abs(got - float(np.sum(decoded))) <= 1e-12 * float(np.sum(np.abs(decoded))), (kind, got, float(np.sum(decoded)))
    part = float(alp_b200.decode_sum(col, first=100, n=237).item())
    want_part = float(xd[100 * 1024 : 337 * 1024].double().sum().item())
    assert abs(part - want_part) <= 1e-9 * max(1.0, abs(want_part))


def test_decimal_sum_of_a_float_column(checker):
    """ALPB200_SUM_DECIMAL (include/alp_b200.h): integers added exactly, one conversion per thread.  Checked against sums made
    on the CHECKER's side: the default path against the float-by-float float64 sum of the checker's decoded column (1e-12
    relative), the decimal path within its stated bound 2^-23 * sum|x| of it — also for a column whose vectors must NOT take
    the decimal path (integers times 10^f beyond int32)."""
    import torch

    import alp_b200
    from oracle import pyoracle

    rng = np.random.default_rng(11)
    cols = {
        "config4": pyoracle.generate(3 * 102400 + 7 * 1024, 4),
        "two_decimals": (rng.integers(-500000, 500000, size=1024 * 150) / 100.0).astype(np.float32),
        "wide_integers": rng.integers(-(1 << 30), 1 << 30, size=1024 * 120).astype(np.float32),
        "near_int32": (rng.integers(0, 2000000, size=1024 * 110) / 1000.0).astype(np.float32),
    }
    for name, x in cols.items():
        col = alp_b200.encode(torch.from_numpy(x).to(_dev()))
        col.read_totals()
        decoded = checker.decode_column(col.to_host())
        assert decoded.tobytes() == x.tobytes(), name
        want = float(np.sum(decoded.astype(np.float64)))
        sum_abs = float(np.sum(np.abs(decoded.astype(np.float64))))
        got = float(alp_b200.decode_sum(col).item())
        assert abs(got - want) <= 1e-12 * sum_abs, (name, got, want)
        dec = float(alp_b200.decode_sum(col, flags=alp_b200.SUM_DECIMAL).item())
        assert abs(dec - want) <= 2.0**-23 * sum_abs, (name, dec, want, sum_abs)
        part = float(alp_b200.decode_sum(col, first=3, n=41, flags=alp_b200.SUM_DECIMAL).item())
        want_part = float(np.sum(decoded[3 * 1024 : 44 * 1024].astype(np.float64)))
        assert abs(part - want_part) <= 2.0**-23 * float(np.sum(np.abs(decoded[3 * 1024 : 44 * 1024].astype(np.float64)))), name
    # doubles accept the flag and ignore it
    xd = alp_b200.generate(102400, 2, _dev())
    cd = alp_b200.encode(xd)
    cd.read_totals()
    a, b = float(alp_b200.decode_sum(cd).item()), float(alp_b200.decode_sum(cd, flags=alp_b200.SUM_DECIMAL).item())
    assert abs(a - b) <= 1e-12 * abs(a)


def test_fused_decode_minmax_count(checker):
    """alpb200_decode_minmax_*: MIN / MAX / COUNT over the decoded values, NaNs ignored — against numpy over the CHECKER's decoded
    column, for ALP, ALP_RD and float columns, exception-heavy vectors, NaN / Inf / -0.0, sub-ranges and the empty range."""
    import torch

    import alp_b200
    from oracle import pyoracle

    rng = np.random.default_rng(21)
    specials = np.round(rng.normal(0, 1000, 1024 * 7), 2)
    specials[rng.integers(0, specials.size, 300)] = np.nan
    specials[5], specials[6000], specials[77] = np.inf, -np.inf, -0.0
    cols = {
        "config2": pyoracle.generate(2 * 102400 + 5 * 1024, 2),
        "config3_rd": pyoracle.generate(102400 + 9 * 1024, 3),
        "config4_f32": pyoracle.generate(102400 + 3 * 1024, 4),
        "negatives": (rng.integers(-10**7, 10**7, size=1024 * 33) / 1000.0),
        "specials": specials,
        "all_nan": np.full(2048, np.nan),
        "random_bits": rng.integers(0, 1 << 63, size=1024 * 11, dtype=np.uint64).view(np.float64),
    }
    for name, x in cols.items():
        col = alp_b200.encode(torch.from_numpy(x).to(_dev()))
        col.read_totals()
        decoded = checker.decode_column(col.to_host()).astype(np.float64)
        for first, n in ((0, col.n_vectors), (1, max(0, col.n_vectors - 2)), (0, 0)):
            part = decoded[first * 1024 : (first + n) * 1024]
            valid = part[~np.isnan(part)]
            mn, mx, cnt = alp_b200.minmax_result(alp_b200.decode_minmax(col, first, n))
            assert cnt == valid.size, (name, first, n, cnt, valid.size)
            if valid.size:
                assert mn == valid.min() and mx == valid.max(), (name, first, n, mn, mx, valid.min(), valid.max())
            else:
                assert mn == np.inf and mx == -np.inf, (name, mn, mx)
    # a stale block-size hint takes the slow path and gives the same answer
    x = pyoracle.generate(102400, 3)
    col = alp_b200.encode(torch.from_numpy(x).to(_dev()))
    col.read_totals()
    col.max_block_bytes = 1024
    mn, mx, cnt = alp_b200.minmax_result(alp_b200.decode_minmax(col))
    assert (mn, mx, cnt) == (float(x.min()), float(x.max()), x.size)


@pytest.mark.parametrize("op", ["<", "<=", ">", ">=", "==", "!="])
def test_fused_decode_filter(op, checker):
    """alpb200_decode_filter_*: the selection bitmap of `value op constant` against numpy over the CHECKER's decoded column
    (ALP, ALP_RD, floats, exception-heavy vectors, NaN / Inf), bit for bit, with the count of selected values."""
    import operator

    import torch

    import alp_b200
    from oracle import pyoracle

    fn = {"<": operator.lt, "<=": operator.le, ">": operator.gt, ">=": operator.ge, "==": operator.eq, "!=": operator.ne}[op]
    rng = np.random.default_rng(31)
    specials = np.round(rng.normal(0, 1000, 1024 * 7), 2)
    specials[rng.integers(0, specials.size, 300)] = np.nan
    specials[5], specials[6000] = np.inf, -np.inf
    cols = {
        "config2": (pyoracle.generate(102400 + 5 * 1024, 2), 123.4),
        "config3_rd": (pyoracle.generate(102400 + 9 * 1024, 3), None),
        "config4_f32": (pyoracle.generate(102400 + 3 * 1024, 4), None),
        "specials": (specials, 0.0),
    }
    for name, (x, c) in cols.items():
        if c is None:
            c = float(np.nanmedian(x.astype(np.float64)))  # a value of the column itself: == / != have something to find
        col = alp_b200.encode(torch.from_numpy(x).to(_dev()))
        col.read_totals()
        decoded = checker.decode_column(col.to_host()).astype(np.float64)
        for first, n in ((0, col.n_vectors), (2, col.n_vectors - 3)):
            bitmap, selected = alp_b200.decode_filter(col, op, c, first, n)
            want = fn(decoded[first * 1024 : (first + n) * 1024], c)
            got = np.unpackbits(bitmap.cpu().numpy().view(np.uint8), bitorder="little").astype(bool)
            assert got.shape == want.shape and np.array_equal(got, want), (name, op, first, n, int(got.sum()), int(want.sum()))
            assert int(selected.item()) == int(want.sum()), (name, op)


def test_fused_decode_sum_propagates_nan():
    import torch

    import alp_b200

    x = torch.arange(4096, dtype=torch.float64, device=_dev()) / 8.0
    x[1234] = float("nan")
    col = alp_b200.encode(x)
    col.read_totals()
    assert torch.isnan(alp_b200.decode_sum(col)).item()


@pytest.mark.parametrize("kind,n_extra", [(2, 0), (2, 777), (3, 0), (4, 5)])
def test_host_codec_sum(kind, n_extra):
    """alpb200_sum_host_*: SUM of a HOST column container (only compressed bytes cross PCIe, one double comes back),
    with and without a padded tail vector; relative 1e-12 against numpy's float64 sum of the original values."""
    import alp_b200
    from oracle import pyoracle

    n = 3 * 102400 + 17 * 1024 + n_extra
    x = pyoracle.generate(n + 1024, kind)[:n].copy()
    codec = alp_b200.HostCodec(n // 1024 + 1, x.dtype.itemsize)
    col = codec.compress(x)
    got = codec.sum(col)
    want = float(np.sum(x.astype(np.float64)))
    assert abs(got - want) <= 1e-12 * float(np.sum(np.abs(x.astype(np.float64)))), (got, want)
    codec.close()


@pytest.mark.parametrize("n_tail", [1, 777, 1023])
def test_host_codec_pads_a_partial_last_vector(n_tail, checker):
    """Columns whose length is not a multiple of 1024: the tail vector is padded on the device, the
    true length travels in the container, and decompress returns exactly the original values."""
    import alp_b200
    from oracle import pyoracle

    n = 102400 + 5 * 1024 + n_tail
    x = pyoracle.generate(n + 1024, 2)[:n].copy()
    codec = alp_b200.HostCodec(n // 1024 + 1, 8)
    col = codec.compress(x)
    assert col.n_vectors == n // 1024 + 1 and col.n_values == n
    assert codec.decompress(col).tobytes() == x.tobytes()
    # the padded column is an ordinary column for everybody else: the checker decodes it; the padding is ONE value of the tail
    # vector (its first non-exception value), and it costs the vector no exception
    full = checker.decode_column(col)
    assert full[:n].tobytes() == x.tobytes()
    pad = full[n:]
    assert np.all(pad == pad[0]) and pad[0] in x[(n // 1024) * 1024 :]
    tail_exc = int(col.meta["exc_cnt"][-1])
    dense = checker.encode_column(full)
    assert int(dense.meta["exc_cnt"][-1]) == tail_exc and col.meta["bw"][-1] == dense.meta["bw"][-1]
    codec.close()


@pytest.mark.parametrize("kind", [2, 3, 4])
@pytest.mark.parametrize("n_tail", [1, 777, 1023])
def test_device_api_tail_vector(kind, n_tail, checker):
    """SURVEY.md section 8f-4 through the DEVICE entry points: alpb200_fill_invalid_* (first strategy before the row-group
    init, second strategy — the first non-exception value — after it), encode, alpb200_decode_values_*.  The filled buffer is
    an ordinary column: the checker, fed the same padded values and the same row-group states, produces the same bytes."""
    import torch

    import alp_b200
    from alp_b200 import _abi

    n = 2 * 102400 + 3 * 1024 + n_tail
    n_vec = -(-n // 1024)
    full = alp_b200.generate(n_vec * 1024, kind, _dev())
    want = full[:n].clone()
    buf = full.clone()
    buf[n:] = float("nan")  # whatever the caller's buffer holds behind its last value
    alp_b200.fill_invalid(buf, n)
    states = alp_b200.rowgroup_init(buf)
    alp_b200.fill_invalid(buf, n, states=states)
    ibits = torch.int64 if buf.element_size() == 8 else torch.int32
    assert torch.equal(buf[:n].view(ibits), want.view(ibits))
    pad = buf[n:]
    assert bool((pad == pad[0]).all()) and bool((buf[(n_vec - 1) * 1024 : n] == pad[0]).any())
    col = alp_b200.encode(buf, states)
    got = alp_b200.decode_values(col, n)
    assert got.numel() == n and torch.equal(got.view(ibits), want.view(ibits))
    # the tail costs no exception beyond those of its real values, and the same bytes come out of the checker
    h = col.to_host()
    host = buf.cpu().numpy()
    st = states.cpu().numpy().view(_abi.RG_STATE_DTYPE).reshape(-1)
    last = host[(n_vec - 1) * 1024 :]
    s_last = st[(n_vec - 1) // 100 : (n_vec - 1) // 100 + 1]
    m = h.meta[-1]
    if int(m["scheme"]) == _abi.SCHEME_ALP:
        ref = checker.encode(last, s_last)
        bw, base = checker.analyze_ffor(ref["enc"])
        assert (ref["cnt"], ref["e"], ref["f"], bw, int(base)) == (int(m["exc_cnt"]), int(m["e"]), int(m["f"]), int(m["bw"]), int(m["base"]))
        assert np.all(ref["pos"] < n - (n_vec - 1) * 1024)  # no exception sits in the padding
        packed = checker.ffor(ref["enc"].view(np.uint64 if buf.element_size() == 8 else np.uint32), bw, base)
        off = int(m["packed_off"]) * 128
        assert h.packed[off : off + 128 * bw].tobytes() == packed.tobytes()
    else:
        ref = checker.rd_encode(last, s_last)
        assert ref["cnt"] == int(m["exc_cnt"]) and np.all(ref["pos"] < n - (n_vec - 1) * 1024)


def test_device_api_nulls():
    """NULL slots (Arrow validity bitmap) are given fillers on the device: a column with 10 % NULLs — whatever garbage the
    caller left in those slots — compresses like the column without them (no extra exceptions, same bit widths), and every
    valid value comes back bit for bit."""
    import torch

    import alp_b200

    n = 3 * 102400
    clean = alp_b200.generate(n, 2, _dev())
    g = torch.Generator(device=_dev()).manual_seed(7)
    valid = torch.rand(n, device=_dev(), generator=g) >= 0.1
    valid[5 * 1024 : 6 * 1024] = False  # one vector without a single valid value
    dirty = clean.clone()
    dirty[~valid] = float("nan")
    bits = valid.view(-1, 8).to(torch.uint8) * (1 << torch.arange(8, device=_dev(), dtype=torch.uint8))
    bitmap = bits.sum(dim=1).to(torch.uint8).contiguous()
    alp_b200.fill_invalid(dirty, n, validity=bitmap)
    assert not bool(torch.isnan(dirty).any())
    states = alp_b200.rowgroup_init(dirty)
    alp_b200.fill_invalid(dirty, n, validity=bitmap, states=states)
    col = alp_b200.encode(dirty, states)
    out = alp_b200.decode(col)
    assert torch.equal(out.view(torch.int64)[valid], clean.view(torch.int64)[valid])
    ref = alp_b200.encode(clean)
    pb, ne = col.read_totals()
    pb_ref, ne_ref = ref.read_totals()
    assert ne <= ne_ref and pb <= pb_ref * 1.001  # the NULLs cost nothing: fewer real values can only mean fewer exceptions
