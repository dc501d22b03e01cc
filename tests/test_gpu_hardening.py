"""GPU tests of the boundary's failure behaviour: stale decode hints, damaged columns, misaligned buffers, concurrent
callers.  The reference's primitives have undefined behaviour on bad input; the C ABI promises error codes
and never touches memory outside the caller's arrays."""
import ctypes
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev():
    import torch

    return torch.device("cuda:0")


def test_stale_block_hint_is_harmless():
    """A decode hint (column.max_block_bytes) smaller than the widest block must not overrun the shared-memory stage:
    the oversized blocks take the direct path and the result is still exact (alp_k_decode.cu:20)."""
    import torch

    import alp_b200

    narrow = alp_b200.generate(100 * 1024, 2, _dev())  # ALP, bw 20: 2560-byte blocks
    wide = alp_b200.generate(100 * 1024, 3, _dev())  # ALP_RD on doubles: 7168-byte blocks
    col_n = alp_b200.encode(narrow)
    col_n.read_totals()
    hint = col_n.max_block_bytes
    assert 0 < hint <= 4096
    # the reuse pattern: a wider column encoded into the same container object
    col = alp_b200.encode(wide, col=col_n)
    assert col.max_block_bytes == 0  # encode() drops the stale hint
    col.max_block_bytes = hint  # a C caller that filled the struct by hand with the old value
    out = alp_b200.decode(col)
    assert torch.equal(out.view(torch.int64), wide.view(torch.int64))
    got = float(alp_b200.decode_sum(col).item())
    want = float(wide.sum().item())
    assert abs(got - want) <= 1e-9 * abs(want) + 1e-6
    # mixed: some blocks fit the stage, some do not
    both = torch.cat([narrow, wide, narrow])
    col2 = alp_b200.encode(both)
    col2.max_block_bytes = 2560
    assert torch.equal(alp_b200.decode(col2).view(torch.int64), both.view(torch.int64))
    # f32
    x4 = alp_b200.generate(64 * 1024, 4, _dev())
    col4 = alp_b200.encode(x4)
    col4.max_block_bytes = 128
    assert torch.equal(alp_b200.decode(col4).view(torch.int32), x4.view(torch.int32))
    s4 = float(alp_b200.decode_sum(col4).item())
    w4 = float(x4.double().sum().item())
    assert abs(s4 - w4) <= 1e-6 * abs(w4)


def _damage(h, rng, how):
    from alp_b200 import _abi

    m = h.meta
    v = int(rng.integers(0, h.n_vectors))
    if how == "bw":
        m["bw"][v] = 200
    elif how == "scheme":
        m["scheme"][v] = 7
    elif how == "exc_cnt":
        m["exc_cnt"][v] = 5000
    elif how == "exc_off":
        m["exc_off"][v] = 0xFFFFFF00
    elif how == "packed_off":
        m["packed_off"][v] = 0xFFFFFF00
    elif how == "e":
        m["e"][v] = 99
    elif how == "rd_left":
        m["scheme"][v] = _abi.SCHEME_ALP_RD
        m["bw"][v] = 3
        m["e"][v] = 200
    elif how == "pos":
        has = np.nonzero(m["exc_cnt"] > 0)[0]
        v = int(has[0])
        h.exc_pos[int(m["exc_off"][v])] = 40000
    return v


@pytest.mark.parametrize("kind", [2, 3, 4])
def test_damaged_columns_are_rejected(kind):
    """A malformed column is rejected with EINVAL by the host entry points (which validate before they launch anything) and by
    alpb200_column_validate_device for device columns; nothing is decoded, nothing faults, and the codec keeps working."""
    import torch

    import alp_b200
    from alp_b200 import _abi

    n_vec = 300
    x = alp_b200.generate(n_vec * 1024, kind, _dev())
    vb = x.element_size()
    good = alp_b200.encode(x).to_host()
    codec = alp_b200.HostCodec(n_vec, vb)
    codec.validate(good)
    dcol = alp_b200.DeviceColumn.from_host(good, _dev()).validate()
    assert dcol.max_block_bytes == int(good.totals[3]) > 0  # validation also renews the decode hint
    rng = np.random.default_rng(kind)
    for how in ("bw", "scheme", "exc_cnt", "exc_off", "packed_off", "e", "rd_left", "pos"):
        h = good.trimmed()
        _damage(h, rng, how)
        with pytest.raises(alp_b200.AlpError) as err:
            codec.validate(h)
        assert err.value.code == _abi.EINVAL, how
        with pytest.raises(alp_b200.AlpError):
            codec.decompress(h)
        with pytest.raises(alp_b200.AlpError):
            codec.sum(h)
        with pytest.raises(alp_b200.AlpError) as err:
            alp_b200.DeviceColumn.from_host(h, _dev()).validate()
        assert err.value.code == _abi.EINVAL, how
    # the codec still works after the rejected calls
    assert codec.decompress(good).tobytes() == x.cpu().numpy().tobytes()
    codec.close()


def test_misaligned_buffers_are_einval():
    import torch

    import alp_b200
    from alp_b200 import _abi

    x = alp_b200.generate(4 * 1024 + 1024, 2, _dev())
    with pytest.raises(alp_b200.AlpError) as err:
        alp_b200.encode(x[1:4097])  # 8-byte aligned only: the bulk-copy engine needs 16
    assert err.value.code == _abi.EINVAL
    col = alp_b200.encode(x[:4096])
    buf = torch.empty(4096 + 2, dtype=torch.float64, device=_dev())
    with pytest.raises(alp_b200.AlpError) as err:
        alp_b200.decode(col, out=buf[1:4097])
    assert err.value.code == _abi.EINVAL
    # nothing sticky: the next good call works
    assert torch.equal(alp_b200.decode(col).view(torch.int64), x[:4096].view(torch.int64))


def test_host_codec_checks_dtype_and_size():
    import alp_b200

    codec = alp_b200.HostCodec(8, 8)
    with pytest.raises(ValueError):
        codec.compress(np.zeros(2048, dtype=np.float32))
    col = codec.compress(np.arange(2048, dtype=np.float64))
    with pytest.raises(ValueError):
        codec.decompress(col, out=np.zeros(1000, dtype=np.float64))
    with pytest.raises(ValueError):
        codec.decompress(col, out=np.zeros(2048, dtype=np.float32))
    assert codec.decompress(col).tolist() == list(range(2048))
    codec.close()


def test_caller_device_is_restored():
    import torch

    import alp_b200

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    torch.cuda.set_device(0)
    codec = alp_b200.HostCodec(8, 8, device=1)
    x = np.arange(4096, dtype=np.float64) / 8
    assert codec.decompress(codec.compress(x)).tobytes() == x.tobytes()
    cur = ctypes.c_int(-1)
    ctypes.CDLL("libcudart.so.12").cudaGetDevice(ctypes.byref(cur))
    assert cur.value == 0
    codec.close()


def test_two_host_threads_two_streams():
    """"Thread-safe for distinct streams" (include/alp_b200.h): two host threads, each with its own torch stream and its
    own codec context, encode / decode / sum different columns at the same time; every result is exact."""
    import torch

    import alp_b200
    from oracle import pyoracle

    errors = []

    def worker(kind, seed, rounds):
        try:
            dev = _dev()
            stream = torch.cuda.Stream(dev)
            n = 6 * 102400
            host = pyoracle.generate(n, kind, seed=seed)
            codec = alp_b200.HostCodec(n // 1024, host.dtype.itemsize)
            with torch.cuda.stream(stream):
                x = torch.from_numpy(host).to(dev)
                ibits = torch.int64 if host.dtype.itemsize == 8 else torch.int32
                for _ in range(rounds):
                    col = alp_b200.encode(x)
                    y = alp_b200.decode(col)
                    s = alp_b200.decode_sum(col)
                    stream.synchronize()
                    assert torch.equal(x.view(ibits), y.view(ibits))
                    want = float(x.double().sum().item())
                    assert abs(float(s.item()) - want) <= 1e-6 * abs(want) + 1e-6
                    h = codec.compress(host)
                    assert codec.decompress(h).tobytes() == host.tobytes()
            codec.close()
        except Exception as exc:  # noqa: BLE001
            errors.append((kind, repr(exc)))

    threads = [threading.Thread(target=worker, args=(k, 100 + i, 6)) for i, k in enumerate((2, 3, 4, 2))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
