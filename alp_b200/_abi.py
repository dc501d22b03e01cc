"""ctypes / numpy images of the structs in include/alp_b200.h.

Pure declarations — no library is loaded here.  The layouts are asserted against the C side by
``alpb200_abi_sizes`` (tests/test_abi.py) and by static_asserts in csrc/alp_capi.cu.
"""
import ctypes

import numpy as np

VECTOR_SIZE = 1024  # reference include/alp/config.hpp:11
ROWGROUP_VECTORS = 100  # config.hpp:13
ROWGROUP_SIZE = VECTOR_SIZE * ROWGROUP_VECTORS  # config.hpp:15
MAX_K = 5  # config.hpp:22
RD_DICT_SIZE = 8  # config.hpp:25
MAX_SAMPLES = 288

SCHEME_INVALID, SCHEME_ALP_RD, SCHEME_ALP = 0, 1, 2  # alp::Scheme, constants.hpp:10-14

OK, EINVAL, ECUDA, ECAPACITY, ENODEVICE = 0, -1, -2, -3, -4

# struct alpb200_rg_state
RG_STATE_DTYPE = np.dtype(
    [
        ("scheme", "<i4"),
        ("k", "<i4"),
        ("combos", "u1", (MAX_K, 2)),
        ("right_bw", "u1"),
        ("left_bw", "u1"),
        ("dict_size", "u1"),
        ("reserved0", "u1", (3,)),
        ("dict", "<u2", (RD_DICT_SIZE,)),
        ("n_extra", "<u2"),
        ("reserved1", "<u2"),
        ("extra_key", "<u2", (MAX_SAMPLES,)),
        ("extra_idx", "<u2", (MAX_SAMPLES,)),
    ]
)
assert RG_STATE_DTYPE.itemsize == 1196

# struct alpb200_vec_meta (32 bytes).  `base` aliases the first 8 bytes of the union; `rd_dict` is exposed as a
# second view of the same 16 bytes by VEC_META_RD_DTYPE.
VEC_META_DTYPE = np.dtype(
    [
        ("base", "<i8"),
        ("reserved_u", "<u8"),
        ("packed_off", "<u4"),
        ("exc_off", "<u4"),
        ("exc_cnt", "<u2"),
        ("scheme", "u1"),
        ("bw", "u1"),
        ("e", "u1"),
        ("f", "u1"),
        ("reserved", "u1", (2,)),
    ]
)
VEC_META_RD_DTYPE = np.dtype(
    [
        ("rd_dict", "<u2", (RD_DICT_SIZE,)),
        ("packed_off", "<u4"),
        ("exc_off", "<u4"),
        ("exc_cnt", "<u2"),
        ("scheme", "u1"),
        ("bw", "u1"),
        ("e", "u1"),
        ("f", "u1"),
        ("reserved", "u1", (2,)),
    ]
)
assert VEC_META_DTYPE.itemsize == 32 and VEC_META_RD_DTYPE.itemsize == 32


class Column(ctypes.Structure):
    """struct alpb200_column"""

    _fields_ = [
        ("n_vectors", ctypes.c_uint64),
        ("meta", ctypes.c_void_p),
        ("packed", ctypes.c_void_p),
        ("packed_capacity", ctypes.c_uint64),
        ("exc_val", ctypes.c_void_p),
        ("exc_pos", ctypes.c_void_p),
        ("exc_capacity", ctypes.c_uint64),
        ("totals", ctypes.c_void_p),
        ("max_block_bytes", ctypes.c_uint64),
        ("n_values", ctypes.c_uint64),
    ]


assert ctypes.sizeof(Column) == 80


def value_types(value_bytes):
    """(float dtype, unsigned dtype, signed dtype, C suffix) for a column of 8- or 4-byte values."""
    if value_bytes == 8:
        return np.dtype("<f8"), np.dtype("<u8"), np.dtype("<i8"), "f64"
    if value_bytes == 4:
        return np.dtype("<f4"), np.dtype("<u4"), np.dtype("<i4"), "f32"
    raise ValueError("value_bytes must be 8 or 4")


def _aligned_empty(nbytes, align=128):
    raw = np.empty(nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off : off + nbytes]


class HostColumn:
    """A column container (struct alpb200_column) whose arrays live in host memory as numpy arrays."""

    def __init__(self, n_vectors, value_bytes, packed_capacity=None, exc_capacity=None):
        self.value_bytes = value_bytes
        self.n_vectors = int(n_vectors)
        if packed_capacity is None:
            packed_capacity = self.n_vectors * VECTOR_SIZE * value_bytes + 128 * 3 * self.n_vectors
        if exc_capacity is None:
            exc_capacity = self.n_vectors * VECTOR_SIZE
        self.meta = np.zeros(self.n_vectors, dtype=VEC_META_DTYPE)
        self.packed = _aligned_empty(int(packed_capacity))
        self.exc_val = np.zeros(int(exc_capacity), dtype=value_types(value_bytes)[1])
        self.exc_pos = np.zeros(int(exc_capacity), dtype=np.uint16)
        self.totals = np.zeros(4, dtype=np.uint64)
        self.n_values = 0  # 0 = n_vectors * 1024 (set by HostCodec.compress when the last vector is padded)

    def as_struct(self):
        return Column(
            self.n_vectors,
            self.meta.ctypes.data,
            self.packed.ctypes.data,
            self.packed.nbytes,
            self.exc_val.ctypes.data,
            self.exc_pos.ctypes.data,
            self.exc_val.shape[0],
            self.totals.ctypes.data,
            int(self.totals[3]),
            int(self.n_values),
        )

    @property
    def packed_bytes(self):
        return int(self.totals[0])

    @property
    def n_exceptions(self):
        return int(self.totals[1])

    def trimmed(self):
        """Copy with arrays cut to the used sizes (what one would store or ship)."""
        out = HostColumn(self.n_vectors, self.value_bytes, max(self.packed_bytes, 128), max(self.n_exceptions, 1))
        out.meta[:] = self.meta
        out.packed[: self.packed_bytes] = self.packed[: self.packed_bytes]
        out.exc_val[: self.n_exceptions] = self.exc_val[: self.n_exceptions]
        out.exc_pos[: self.n_exceptions] = self.exc_pos[: self.n_exceptions]
        out.totals[:] = self.totals
        return out

    def compressed_bytes(self):
        """Algorithmic size of the compressed column: packed + 13/9-byte headers + exceptions."""
        hdr = 5 + self.value_bytes
        alp = self.meta["scheme"] == SCHEME_ALP
        exc = self.meta["exc_cnt"].astype(np.int64)
        body = self.packed_bytes
        alp_bytes = int((hdr + exc[alp] * (self.value_bytes + 2)).sum())
        rd_bytes = int((4 + exc[~alp] * 4).sum())
        return body + alp_bytes + rd_bytes
