"""The reference's per-vector primitive API (cwida/ALP PRIMITIVES.md), one 1024-value vector per call, numpy in/out.

Each function is a thin wrapper over an ``alpb200_prim_*`` entry point of the C ABI: a 1-vector batch that is
copied to the GPU, processed by the same device code as the batched kernels, and copied back.  The names and
argument meaning follow the reference:

================================  =====================================================================
here                              reference
================================  =====================================================================
init(col, offset)                 alp::encoder<PT>::init (+ rd_encoder<PT>::init)   encoder.hpp:420, rd.hpp:180
encode(vec, state)                alp::encoder<PT>::encode                           encoder.hpp:402
analyze_ffor(enc)                 alp::encoder<PT>::analyze_ffor                     encoder.hpp:109
ffor(values, bw, base)            ffor::ffor                                         fastlanes/ffor.hpp:7-15
unffor(packed, bw, base, dtype)   unffor::unffor                                     fastlanes/unffor.hpp:7-15
falp(packed, bw, base, f, e)      generated::falp::fallback::scalar::falp            alp/falp.hpp:10-44
decode(enc, f, e)                 alp::decoder<PT>::decode                           decoder.hpp:134
patch(out, exc, pos)              alp::decoder<PT>::patch_exceptions                 decoder.hpp:141
rd_encode(vec, state)             alp::rd_encoder<PT>::encode                        rd.hpp:109
rd_decode(right, left, ...)       alp::rd_encoder<PT>::decode                        rd.hpp:152
================================  =====================================================================

The function set is deliberately the same as the CPU checkers' numpy driver, so a test can run the GPU and a
checker through identical code.
"""
import ctypes

import numpy as np

from . import _abi
from ._lib import check, lib

kind = "gpu"


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def init(col, offset=0):
    col = np.ascontiguousarray(col)
    sfx = _abi.value_types(col.dtype.itemsize)[3]
    st = np.zeros(1, dtype=_abi.RG_STATE_DTYPE)
    check(getattr(lib, "alpb200_prim_init_" + sfx)(_p(col), offset, col.shape[0], _p(st)))
    return st


def encode(vec, state):
    vec = np.ascontiguousarray(vec)
    ft, ut, it, sfx = _abi.value_types(vec.dtype.itemsize)
    exc = np.zeros(1024, dtype=ft)
    pos = np.zeros(1024, dtype=np.uint16)
    cnt = np.zeros(1, dtype=np.uint16)
    enc = np.zeros(1024, dtype=it)
    ef = np.zeros(2, dtype=np.uint8)
    state = np.ascontiguousarray(state)
    check(getattr(lib, "alpb200_prim_encode_" + sfx)(_p(vec), _p(state), _p(exc), _p(pos), _p(cnt), _p(enc), _p(ef[0:1]), _p(ef[1:2])))
    n = int(cnt[0])
    return dict(enc=enc, exc=exc[:n].copy(), pos=pos[:n].copy(), cnt=n, e=int(ef[0]), f=int(ef[1]))


def analyze_ffor(enc):
    enc = np.ascontiguousarray(enc)
    bw = np.zeros(1, dtype=np.uint8)
    base = np.zeros(1, dtype=enc.dtype)
    fn = lib.alpb200_prim_analyze_ffor_i64 if enc.dtype.itemsize == 8 else lib.alpb200_prim_analyze_ffor_i32
    check(fn(_p(enc), _p(bw), _p(base)))
    return int(bw[0]), base[0]


def ffor(values, bw, base=0):
    values = np.ascontiguousarray(values)
    t = values.dtype.itemsize * 8
    out = np.zeros(1024, dtype=values.dtype)
    check(getattr(lib, "alpb200_prim_ffor_u%d" % t)(_p(values), _p(out), bw, int(base) & ((1 << t) - 1)))
    return out[: bw * 1024 // t].copy()


def unffor(packed, bw, base=0, dtype=None):
    dtype = np.dtype(dtype or packed.dtype)
    t = dtype.itemsize * 8
    buf = np.zeros(1024, dtype=dtype)
    buf[: packed.shape[0]] = packed
    out = np.zeros(1024, dtype=dtype)
    check(getattr(lib, "alpb200_prim_unffor_u%d" % t)(_p(buf), _p(out), bw, int(base) & ((1 << t) - 1)))
    return out


def falp(packed, bw, base, f, e, value_bytes=8):
    ft, ut, it, sfx = _abi.value_types(value_bytes)
    buf = np.zeros(1024, dtype=ut)
    buf[: packed.shape[0]] = packed
    out = np.zeros(1024, dtype=ft)
    check(getattr(lib, "alpb200_prim_falp_" + sfx)(_p(buf), _p(out), bw, int(base) & ((1 << (8 * value_bytes)) - 1), f, e))
    return out


def decode(enc, f, e):
    enc = np.ascontiguousarray(enc)
    ft, ut, it, sfx = _abi.value_types(enc.dtype.itemsize)
    out = np.zeros(1024, dtype=ft)
    check(getattr(lib, "alpb200_prim_decode_" + sfx)(_p(enc), f, e, _p(out)))
    return out


def patch(out, exc, pos):
    sfx = _abi.value_types(out.dtype.itemsize)[3]
    exc = np.ascontiguousarray(exc, dtype=out.dtype)
    pos = np.ascontiguousarray(pos, dtype=np.uint16)
    check(getattr(lib, "alpb200_prim_patch_" + sfx)(_p(out), _p(exc), _p(pos), len(pos)))
    return out


def rd_encode(vec, state):
    vec = np.ascontiguousarray(vec)
    ft, ut, it, sfx = _abi.value_types(vec.dtype.itemsize)
    exc = np.zeros(1024, dtype=np.uint16)
    pos = np.zeros(1024, dtype=np.uint16)
    cnt = np.zeros(1, dtype=np.uint16)
    right = np.zeros(1024, dtype=ut)
    left = np.zeros(1024, dtype=np.uint16)
    state = np.ascontiguousarray(state)
    check(getattr(lib, "alpb200_prim_rd_encode_" + sfx)(_p(vec), _p(state), _p(exc), _p(pos), _p(cnt), _p(right), _p(left)))
    n = int(cnt[0])
    return dict(right=right, left=left, exc=exc[:n].copy(), pos=pos[:n].copy(), cnt=n)


def rd_decode(right, left, exc, pos, state):
    right = np.ascontiguousarray(right)
    ft, ut, it, sfx = _abi.value_types(right.dtype.itemsize)
    out = np.zeros(1024, dtype=ft)
    exc = np.ascontiguousarray(exc, dtype=np.uint16)
    pos = np.ascontiguousarray(pos, dtype=np.uint16)
    left = np.ascontiguousarray(left, dtype=np.uint16)
    state = np.ascontiguousarray(state)
    check(getattr(lib, "alpb200_prim_rd_decode_" + sfx)(_p(out), _p(right), _p(left), _p(exc), _p(pos), len(pos), _p(state)))
    return out
