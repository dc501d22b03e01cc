"""Batched column codec on the device, and the host-buffer codec context.

This is the layer the reference leaves to its callers (the row-group loop of benchmarks/benchmark.cpp:200-285 and
publication/source_code/bench_compression_ratio/alp.cpp:198-229): ``rowgroup_init`` once per 100 vectors, then
``encode`` / ``decode`` for every vector — here as one kernel launch over the whole column.

torch tensors are only the owners of device memory; all work is done by libalp_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _abi
from ._lib import check, lib

_FLOAT = {8: torch.float64, 4: torch.float32}


def device_count():
    """CUDA devices the library can see (raises AlpError without one: there is no CPU fallback)."""
    return check(lib.alpb200_device_count())


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or not t.is_contiguous():
        raise ValueError("%s must be a contiguous CUDA tensor" % what)


def _sfx(value_bytes):
    return _abi.value_types(value_bytes)[3]


def worst_case_capacities(n_vectors, value_bytes):
    """(packed bytes, exception slots) that can never overflow: every vector ALP_RD at full width, all exceptions."""
    units = 66 if value_bytes == 8 else 35
    return n_vectors * units * 128, n_vectors * _abi.VECTOR_SIZE


class DeviceColumn:
    """struct alpb200_column whose arrays are CUDA tensors."""

    def __init__(self, n_vectors, value_bytes, device, packed_capacity=None, exc_capacity=None):
        wp, we = worst_case_capacities(n_vectors, value_bytes)
        self.n_vectors = int(n_vectors)
        self.value_bytes = value_bytes
        self.device = torch.device(device)
        self.packed_capacity = int(wp if packed_capacity is None else packed_capacity)
        self.exc_capacity = int(we if exc_capacity is None else exc_capacity)
        self.meta = torch.zeros((self.n_vectors, 32), dtype=torch.uint8, device=self.device)
        self.packed = torch.empty(max(self.packed_capacity, 128), dtype=torch.uint8, device=self.device)
        self.exc_val = torch.empty(max(self.exc_capacity, 1), dtype=torch.int64 if value_bytes == 8 else torch.int32, device=self.device)
        self.exc_pos = torch.empty(max(self.exc_capacity, 1), dtype=torch.int16, device=self.device)
        self.totals = torch.zeros(4, dtype=torch.int64, device=self.device)
        self.max_block_bytes = 0
        assert self.packed.data_ptr() % 128 == 0

    def as_struct(self):
        return _abi.Column(
            self.n_vectors,
            self.meta.data_ptr(),
            self.packed.data_ptr(),
            self.packed_capacity,
            self.exc_val.data_ptr(),
            self.exc_pos.data_ptr(),
            self.exc_capacity,
            self.totals.data_ptr(),
            self.max_block_bytes,
            0,
        )

    def read_totals(self):
        """Synchronising read of (packed bytes, exceptions, overflow flag, widest block); updates max_block_bytes."""
        t = self.totals.cpu().numpy().astype(np.uint64)
        if int(t[2]) != 0:
            raise RuntimeError("alp_b200: the column container overflowed during encode (capacities too small)")
        self.max_block_bytes = int(t[3])
        return int(t[0]), int(t[1])

    def to_host(self, first=0, n=None):
        """Vectors [first, first+n) as a HostColumn of their own (offsets rebased to the slice)."""
        self.read_totals()
        n = self.n_vectors - first if n is None else n
        meta = self.meta[first : first + n].cpu().numpy().view(_abi.VEC_META_DTYPE).reshape(-1).copy()
        if n == 0:
            return _abi.HostColumn(0, self.value_bytes, 128, 1)
        units = np.where(meta["scheme"] == _abi.SCHEME_ALP_RD, meta["bw"].astype(np.int64) + meta["e"], meta["bw"].astype(np.int64))
        # min / max over the records: exact for a vector-order column, a slight superset for a completion-order one
        p0, e0 = int(meta["packed_off"].min()) * 128, int(meta["exc_off"].min())
        p1 = int((meta["packed_off"].astype(np.int64) + units).max()) * 128
        e1 = int((meta["exc_off"].astype(np.int64) + meta["exc_cnt"]).max())
        h = _abi.HostColumn(n, self.value_bytes, max(p1 - p0, 128), max(e1 - e0, 1))
        meta["packed_off"] -= p0 // 128
        meta["exc_off"] -= e0
        h.meta[:] = meta
        h.packed[: p1 - p0] = self.packed[p0:p1].cpu().numpy()
        h.exc_val[: e1 - e0] = self.exc_val[e0:e1].cpu().numpy().view(h.exc_val.dtype)
        h.exc_pos[: e1 - e0] = self.exc_pos[e0:e1].cpu().numpy().view(np.uint16)
        h.totals[:] = [p1 - p0, e1 - e0, 0, self.max_block_bytes]
        return h

    @classmethod
    def from_host(cls, h, device):
        packed_bytes, n_exc = h.packed_bytes, h.n_exceptions
        col = cls(h.n_vectors, h.value_bytes, device, max(packed_bytes, 128), max(n_exc, 1))
        col.meta.copy_(torch.from_numpy(h.meta.view(np.uint8).reshape(-1, 32)))
        col.packed[:packed_bytes].copy_(torch.from_numpy(np.ascontiguousarray(h.packed[:packed_bytes])))
        signed = np.int64 if h.value_bytes == 8 else np.int32
        col.exc_val[:n_exc].copy_(torch.from_numpy(h.exc_val[:n_exc].view(signed)))
        col.exc_pos[:n_exc].copy_(torch.from_numpy(h.exc_pos[:n_exc].view(np.int16)))
        col.totals.copy_(torch.tensor([packed_bytes, n_exc, 0, int(h.totals[3])], dtype=torch.int64))
        col.max_block_bytes = int(h.totals[3])
        return col

    def validate(self):
        """alpb200_column_validate_device: raises AlpError(EINVAL) for a malformed column; on success also renews the
        decode hint (max_block_bytes) from the records.  Synchronises the current stream."""
        st = self.as_struct()
        widest = ctypes.c_uint64(0)
        with torch.cuda.device(self.device):
            check(lib.alpb200_column_validate_device(ctypes.byref(st), self.value_bytes, ctypes.byref(widest), _stream_ptr(self.device)))
        self.max_block_bytes = int(widest.value)
        return self

    def shard(self, first, n):
        """A view of vectors [first, first+n) as a column of its own (metadata copied and rebased on the host side
        is not needed: offsets stay absolute and the arrays are shared)."""
        view = object.__new__(DeviceColumn)
        view.__dict__.update(self.__dict__)
        view.n_vectors = n
        view.meta = self.meta[first : first + n]
        return view


def rowgroup_init(values, states=None, workspace=None):
    """alp::encoder<PT>::init (+ rd_encoder<PT>::init) for every row-group of a device column → uint8 [n_rg, 1196].
    `states` / `workspace` let a caller reuse its buffers (nothing is allocated then)."""
    _require_cuda(values, "values")
    vb = values.element_size()
    n = values.numel()
    n_rg = max(1, -(-(n // _abi.VECTOR_SIZE) // _abi.ROWGROUP_VECTORS))
    if states is None:
        states = torch.empty((n_rg, _abi.RG_STATE_DTYPE.itemsize), dtype=torch.uint8, device=values.device)
    _require_cuda(states, "states")
    if states.numel() < n_rg * _abi.RG_STATE_DTYPE.itemsize:
        raise ValueError("states holds fewer than %d row-group records" % n_rg)
    need = max(256, lib.alpb200_init_workspace_bytes(n))
    ws = torch.empty(need, dtype=torch.uint8, device=values.device) if workspace is None else workspace
    if ws.numel() < need:
        raise ValueError("workspace is smaller than alpb200_init_workspace_bytes(n)")
    with torch.cuda.device(values.device):
        fn = getattr(lib, "alpb200_rowgroup_init_" + _sfx(vb))
        check(fn(values.data_ptr(), n, states.data_ptr(), ws.data_ptr(), _stream_ptr(values.device)))
    return states


def encode(values, states=None, col=None, workspace=None, ordered=True, append_at=None):
    """Compress a device column (numel a multiple of 1024) → DeviceColumn.  `states` defaults to rowgroup_init(values).
    ordered=False selects the completion-order layout (alpb200_encode_unordered_*): same blocks, faster, bytes not reproducible.
    append_at=v (a multiple of 100, with `col` given): `values` are the vectors v, v+1, ... of `col`, whose earlier vectors
    were encoded by previous calls; the output continues where the column ends (ALPB200_ENCODE_APPEND)."""
    _require_cuda(values, "values")
    vb = values.element_size()
    if values.dtype != _FLOAT[vb] or values.numel() % _abi.VECTOR_SIZE:
        raise ValueError("values must be float64/float32 with a multiple of 1024 elements")
    n_vec = values.numel() // _abi.VECTOR_SIZE
    if states is None:
        states = rowgroup_init(values)
    _require_cuda(states, "states")
    if col is None:
        if append_at is not None:
            raise ValueError("append_at needs the column that is being continued")
        col = DeviceColumn(n_vec, vb, values.device)
    if workspace is None:
        workspace = torch.empty(max(256, lib.alpb200_encode_workspace_bytes(n_vec)), dtype=torch.uint8, device=values.device)
    col.max_block_bytes = 0  # the decode hint of whatever this column held before is stale now (read_totals() renews it)
    st = col.as_struct()
    flags = 0 if ordered else 1
    if append_at is not None:
        if append_at % _abi.ROWGROUP_VECTORS or append_at + n_vec > col.n_vectors:
            raise ValueError("append_at must start a row-group inside the column")
        st.meta = col.meta.data_ptr() + append_at * 32
        st.n_vectors = n_vec
        flags |= 2 if append_at > 0 else 0
    with torch.cuda.device(values.device):
        fn = getattr(lib, "alpb200_encode_ex_" + _sfx(vb))
        check(fn(values.data_ptr(), n_vec, states.data_ptr(), ctypes.byref(st), workspace.data_ptr(), _stream_ptr(values.device), flags))
    return col


def decode(col, first=0, n=None, out=None):
    """Decompress vectors [first, first+n) of a DeviceColumn → tensor of n*1024 values."""
    n = col.n_vectors - first if n is None else n
    if out is None:
        out = torch.empty(n * _abi.VECTOR_SIZE, dtype=_FLOAT[col.value_bytes], device=col.device)
    _require_cuda(out, "out")
    st = col.as_struct()
    with torch.cuda.device(col.device):
        fn = getattr(lib, "alpb200_decode_" + _sfx(col.value_bytes))
        check(fn(ctypes.byref(st), first, n, out.data_ptr(), _stream_ptr(col.device)))
    return out


def fill_invalid(values, n_values, validity=None, states=None):
    """alpb200_fill_invalid_*: fillers for NULL slots (validity: uint8 CUDA tensor, Arrow bitmap, bit set = valid) and for the
    slots behind n_values; `values` must hold ceil(n_values / 1024) * 1024 elements.  states=None: first valid value of the
    vector (run before rowgroup_init); states given: first valid non-exception value (run between init and encode)."""
    _require_cuda(values, "values")
    vb = values.element_size()
    if values.numel() < -(-n_values // _abi.VECTOR_SIZE) * _abi.VECTOR_SIZE:
        raise ValueError("values must have room for whole vectors")
    with torch.cuda.device(values.device):
        fn = getattr(lib, "alpb200_fill_invalid_" + _sfx(vb))
        check(fn(values.data_ptr(), n_values, None if validity is None else validity.data_ptr(), None if states is None else states.data_ptr(),
                 _stream_ptr(values.device)))
    return values


def decode_values(col, n_values, first=0, out=None):
    """alpb200_decode_values_*: exactly n_values values from vector `first` on (a partial last vector goes through scratch)."""
    if out is None:
        out = torch.empty(n_values, dtype=_FLOAT[col.value_bytes], device=col.device)
    _require_cuda(out, "out")
    scratch = torch.empty(_abi.VECTOR_SIZE, dtype=_FLOAT[col.value_bytes], device=col.device)
    st = col.as_struct()
    with torch.cuda.device(col.device):
        fn = getattr(lib, "alpb200_decode_values_" + _sfx(col.value_bytes))
        check(fn(ctypes.byref(st), first, n_values, out.data_ptr(), scratch.data_ptr(), _stream_ptr(col.device)))
    return out


SUM_DECIMAL = 1  # ALPB200_SUM_DECIMAL


def decode_sum(col, first=0, n=None, out=None, flags=0):
    """Fused decode + SUM of vectors [first, first+n): adds into the 1-element float64 CUDA tensor `out` (created
    zeroed when omitted).  Nothing is written back to HBM; addition order is not fixed.  flags: SUM_DECIMAL (float columns:
    add the integers and convert once, see include/alp_b200.h)."""
    n = col.n_vectors - first if n is None else n
    if out is None:
        out = torch.zeros(1, dtype=torch.float64, device=col.device)
    _require_cuda(out, "out")
    st = col.as_struct()
    with torch.cuda.device(col.device):
        fn = getattr(lib, "alpb200_decode_sum_ex_" + _sfx(col.value_bytes))
        check(fn(ctypes.byref(st), first, n, out.data_ptr(), flags, _stream_ptr(col.device)))
    return out


def decode_minmax(col, first=0, n=None, out=None):
    """Fused decode + MIN / MAX / COUNT of vectors [first, first+n): returns the 3-element float64 CUDA tensor `out`
    (min, max, count-as-bits); use minmax_result() to read it.  NaNs are ignored; nothing is written back to HBM."""
    n = col.n_vectors - first if n is None else n
    if out is None:
        out = torch.empty(3, dtype=torch.float64, device=col.device)
    _require_cuda(out, "out")
    st = col.as_struct()
    with torch.cuda.device(col.device):
        fn = getattr(lib, "alpb200_decode_minmax_" + _sfx(col.value_bytes))
        check(fn(ctypes.byref(st), first, n, out.data_ptr(), _stream_ptr(col.device)))
    return out


FILTER_OPS = {"<": 0, "<=": 1, ">": 2, ">=": 3, "==": 4, "!=": 5}  # ALPB200_FILTER_*


def decode_filter(col, op, constant, first=0, n=None, bitmap=None, selected=None):
    """Fused decode + predicate filter of vectors [first, first+n): returns (bitmap, selected) — an int32 CUDA tensor of
    32 words per vector (bit j of word w of a vector = its value 32 w + j satisfies `value op constant`) and a 1-element
    int64 CUDA tensor with the number of set bits.  op: one of FILTER_OPS' keys."""
    n = col.n_vectors - first if n is None else n
    if bitmap is None:
        bitmap = torch.empty(32 * n, dtype=torch.int32, device=col.device)
    if selected is None:
        selected = torch.zeros(1, dtype=torch.int64, device=col.device)
    _require_cuda(bitmap, "bitmap")
    st = col.as_struct()
    with torch.cuda.device(col.device):
        fn = getattr(lib, "alpb200_decode_filter_" + _sfx(col.value_bytes))
        check(fn(ctypes.byref(st), first, n, FILTER_OPS[op], float(constant), bitmap.data_ptr(), selected.data_ptr(), _stream_ptr(col.device)))
    return bitmap, selected


def minmax_result(out):
    """(min, max, count) from decode_minmax's tensor (synchronises)."""
    h = out.cpu()
    return float(h[0]), float(h[1]), int(h.view(torch.int64)[2])


def generate(n_values, kind, device, seed=None, first_index=0, out=None):
    """Synthetic columns of SURVEY.md §8d on the device: kind 2 decimal f64, 3 high-precision f64, 4 mixed f32."""
    seed = {2: 42, 3: 43, 4: 44}[kind] if seed is None else seed
    dtype = torch.float32 if kind == 4 else torch.float64
    if out is None:
        out = torch.empty(n_values, dtype=dtype, device=device)
    with torch.cuda.device(out.device):
        fn = lib.alpb200_generate_f32 if kind == 4 else lib.alpb200_generate_f64
        check(fn(out.data_ptr(), n_values, first_index, seed, kind, _stream_ptr(out.device)))
    return out


class HostCodec:
    """alpb200_ctx: compress / decompress columns that live in HOST memory (copies are part of each call)."""

    OPT_UNORDERED = 1  # ALPB200_OPT_UNORDERED
    OPT_CHUNKS = 2  # ALPB200_OPT_CHUNKS

    def __init__(self, max_vectors, value_bytes=8, device=0, ordered=True, packed_capacity=0, exc_capacity=0, chunks=None):
        """packed_capacity / exc_capacity: the caller's bound on the compressed column (bytes / exception slots) the device
        staging is sized for; 0 = the worst case any column can reach (~27 bytes per f64 value)."""
        self.value_bytes = value_bytes
        self.max_vectors = int(max_vectors)
        self._ctx = ctypes.c_void_p()
        check(lib.alpb200_ctx_create_ex(ctypes.byref(self._ctx), device, self.max_vectors, value_bytes, int(packed_capacity), int(exc_capacity)))
        if not ordered:
            check(lib.alpb200_ctx_set_option(self._ctx, self.OPT_UNORDERED, 1))
        if chunks is not None:
            check(lib.alpb200_ctx_set_option(self._ctx, self.OPT_CHUNKS, int(chunks)))

    def close(self):
        if self._ctx:
            lib.alpb200_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check_values(self, arr, what, n_min=0):
        want = _abi.value_types(self.value_bytes)[0]
        if not isinstance(arr, np.ndarray) or arr.dtype != want or arr.ndim != 1 or not arr.flags["C_CONTIGUOUS"]:
            raise ValueError("%s must be a contiguous 1-D numpy array of %s" % (what, want))
        if arr.shape[0] < n_min:
            raise ValueError("%s holds %d values, %d are needed" % (what, arr.shape[0], n_min))

    def compress(self, values, col=None):
        values = np.ascontiguousarray(values)
        self._check_values(values, "values")
        n_vec = -(-values.shape[0] // _abi.VECTOR_SIZE)  # a partial last vector is padded by the library
        if col is None:
            col = _abi.HostColumn(n_vec, self.value_bytes)
        st = col.as_struct()
        fn = getattr(lib, "alpb200_compress_host_" + _sfx(self.value_bytes))
        check(fn(self._ctx, values.ctypes.data, values.shape[0], ctypes.byref(st)))
        col.n_vectors = int(st.n_vectors)
        col.n_values = int(st.n_values)
        return col

    def sum(self, col):
        """SUM of a host column: only the compressed bytes go to the device, one double comes back (alpb200_sum_host_*)."""
        st = col.as_struct()
        out = ctypes.c_double(0.0)
        fn = getattr(lib, "alpb200_sum_host_" + _sfx(self.value_bytes))
        check(fn(self._ctx, ctypes.byref(st), ctypes.byref(out)))
        return float(out.value)

    def validate(self, col):
        """alpb200_column_validate_host: raises AlpError(EINVAL) for a malformed host column."""
        st = col.as_struct()
        check(lib.alpb200_column_validate_host(ctypes.byref(st), self.value_bytes))

    def decompress(self, col, out=None):
        n_out = col.n_values or col.n_vectors * _abi.VECTOR_SIZE
        if col.value_bytes != self.value_bytes:
            raise ValueError("the column holds %d-byte values, the codec was created for %d" % (col.value_bytes, self.value_bytes))
        if out is None:
            out = np.empty(n_out, dtype=_abi.value_types(self.value_bytes)[0])
        self._check_values(out, "out", n_out)
        st = col.as_struct()
        fn = getattr(lib, "alpb200_decompress_host_" + _sfx(self.value_bytes))
        check(fn(self._ctx, ctypes.byref(st), out.ctypes.data))
        return out
