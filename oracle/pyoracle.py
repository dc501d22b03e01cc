"""ctypes driver for the CPU checkers — TEST INFRASTRUCTURE, not product code.

Two libraries with the same function set (see oracle/alp_oracle.h and oracle/ref_shim.cpp):

* ``port()``       — oracle/liboracle.so, the plain-C restatement (prefix ``alpo_``)
* ``reference()``  — oracle/_ref/libalp_ref_v{3,4}.so, the unmodified reference compiled from
                     /root/reference (prefix ``alpref_``); ``None`` when it was never built

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import this.
"""
import ctypes
import importlib.util
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _load_abi():
    """The struct declarations of include/alp_b200.h (alp_b200/_abi.py: pure ctypes / numpy images, no library) WITHOUT
    importing the alp_b200 package, whose __init__ loads the CUDA library: a process that only runs the CPU checker
    (bench.py --impl reference) must not map the product."""
    mod = sys.modules.get("alp_b200._abi")
    if mod is None:
        spec = importlib.util.spec_from_file_location("alp_b200._abi", os.path.join(os.path.dirname(_HERE), "alp_b200", "_abi.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["alp_b200._abi"] = mod  # a later `import alp_b200` picks up this very module
        spec.loader.exec_module(mod)
    return mod


_abi = _load_abi()
_c = ctypes
_P = ctypes.c_void_p


def build(verbose=False):
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref (building the checker is not using it)."""
    res = subprocess.run(["make", "-C", _HERE, "-j8", "all"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("oracle build failed")


def _cpu_has_avx512():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    flags = set(line.split(":")[1].split())
                    return {"avx512f", "avx512dq", "avx512bw", "avx512vl", "avx512cd"} <= flags
    except OSError:
        pass
    return False


def _ptr(a):
    return a.ctypes.data_as(_P)


class CpuCodec:
    """numpy face of one checker library."""

    def __init__(self, path, prefix, kind):
        self.lib = ctypes.CDLL(path)
        self.path = path
        self.prefix = prefix
        self.kind = kind  # "port" | "reference"
        fn = getattr(self.lib, prefix + "build_info")
        fn.restype = ctypes.c_char_p
        self.build_info = fn().decode()
        sizes = (ctypes.c_uint32 * 3)()
        getattr(self.lib, prefix + "abi_sizes")(sizes)
        want = (_abi.RG_STATE_DTYPE.itemsize, _abi.VEC_META_DTYPE.itemsize, ctypes.sizeof(_abi.Column))
        if tuple(sizes) != want:
            raise RuntimeError("%s was built against a different include/alp_b200.h (%s != %s): rebuild it" % (path, tuple(sizes), want))
        for sfx, ct in (("f64", _c.c_double), ("f32", _c.c_float)):
            it = _c.c_int64 if sfx == "f64" else _c.c_int32
            f = self._fn("encode_value_" + sfx)
            f.argtypes, f.restype = [ct, _c.c_uint8, _c.c_uint8], it
            f = self._fn("decode_value_" + sfx)
            f.argtypes, f.restype = [it, _c.c_uint8, _c.c_uint8], ct
        for t, ct in ((64, _c.c_uint64), (32, _c.c_uint32), (16, _c.c_uint16), (8, _c.c_uint8)):
            for name in ("ffor", "unffor"):
                self._fn("%s_u%d" % (name, t)).argtypes = [_P, _P, _c.c_uint8, ct]
        self._fn("falp_f64").argtypes = [_P, _P, _c.c_uint8, _c.c_uint64, _c.c_uint8, _c.c_uint8]
        self._fn("falp_f32").argtypes = [_P, _P, _c.c_uint8, _c.c_uint32, _c.c_uint8, _c.c_uint8]
        for sfx in ("f64", "f32"):
            self._fn("init_" + sfx).argtypes = [_P, _c.c_size_t, _c.c_size_t, _P]
            self._fn("encode_" + sfx).argtypes = [_P] * 8
            self._fn("decode_" + sfx).argtypes = [_P, _c.c_uint8, _c.c_uint8, _P]
            self._fn("patch_" + sfx).argtypes = [_P, _P, _P, _c.c_uint16]
            self._fn("rd_encode_" + sfx).argtypes = [_P] * 7
            self._fn("rd_decode_" + sfx).argtypes = [_P, _P, _P, _P, _P, _c.c_uint16, _P]
            f = self._fn("encode_column_" + sfx)
            f.argtypes, f.restype = [_P, _c.c_size_t, _c.c_int, _P], _c.c_int
            f = self._fn("decode_column_" + sfx)
            f.argtypes, f.restype = [_P, _c.c_size_t, _c.c_size_t, _c.c_int, _P], _c.c_int
            f = self._fn("bench_encode_" + sfx)
            f.argtypes, f.restype = [_P, _c.c_size_t, _c.c_int, _P, _P, _c.c_int], _c.c_int
            f = self._fn("sum_column_" + sfx)
            f.argtypes, f.restype = [_P, _c.c_size_t, _c.c_size_t, _c.c_int, _P], _c.c_int
        self._fn("analyze_ffor_i64").argtypes = [_P, _P, _P]
        self._fn("analyze_ffor_i32").argtypes = [_P, _P, _P]

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    # -- scalar helpers -------------------------------------------------------------------------
    def encode_value(self, v, f, e, value_bytes=8):
        sfx = _abi.value_types(value_bytes)[3]
        return self._fn("encode_value_" + sfx)(v, f, e)

    def decode_value(self, x, f, e, value_bytes=8):
        sfx = _abi.value_types(value_bytes)[3]
        return self._fn("decode_value_" + sfx)(x, f, e)

    # -- row-group init -------------------------------------------------------------------------
    def init(self, col, offset=0):
        """encoder<PT>::init (+ rd_encoder<PT>::init) for the row-group starting at `offset` → RG_STATE record."""
        col = np.ascontiguousarray(col)
        sfx = _abi.value_types(col.dtype.itemsize)[3]
        st = np.zeros(1, dtype=_abi.RG_STATE_DTYPE)
        self._fn("init_" + sfx)(_ptr(col), offset, col.shape[0], _ptr(st))
        return st

    # -- ALP vector primitives --------------------------------------------------------------------
    def encode(self, vec, state):
        """encoder<PT>::encode → dict(enc, exc, pos, cnt, e, f)"""
        vec = np.ascontiguousarray(vec)
        ft, ut, it, sfx = _abi.value_types(vec.dtype.itemsize)
        exc = np.zeros(1024, dtype=ft)
        pos = np.zeros(1024, dtype=np.uint16)
        cnt = np.zeros(1, dtype=np.uint16)
        enc = np.zeros(1024, dtype=it)
        ef = np.zeros(2, dtype=np.uint8)
        self._fn("encode_" + sfx)(
            _ptr(vec), _ptr(state), _ptr(exc), _ptr(pos), _ptr(cnt), _ptr(enc), _ptr(ef[0:1]), _ptr(ef[1:2])
        )
        n = int(cnt[0])
        return dict(enc=enc, exc=exc[:n].copy(), pos=pos[:n].copy(), cnt=n, e=int(ef[0]), f=int(ef[1]))

    def analyze_ffor(self, enc):
        enc = np.ascontiguousarray(enc)
        bw = np.zeros(1, dtype=np.uint8)
        base = np.zeros(1, dtype=enc.dtype)
        self._fn("analyze_ffor_i64" if enc.dtype.itemsize == 8 else "analyze_ffor_i32")(_ptr(enc), _ptr(bw), _ptr(base))
        return int(bw[0]), base[0]

    def ffor(self, values, bw, base=0):
        """ffor::ffor on unsigned lanes of width values.dtype → packed words (bw*1024/T of them)."""
        values = np.ascontiguousarray(values)
        t = values.dtype.itemsize * 8
        out = np.zeros(1024, dtype=values.dtype)
        self._fn("ffor_u%d" % t)(_ptr(values), _ptr(out), bw, int(base) & ((1 << t) - 1))
        return out[: bw * 1024 // t].copy()

    def unffor(self, packed, bw, base=0, dtype=None):
        dtype = np.dtype(dtype or packed.dtype)
        t = dtype.itemsize * 8
        buf = np.zeros(1024, dtype=dtype)
        buf[: packed.shape[0]] = packed
        out = np.zeros(1024, dtype=dtype)
        self._fn("unffor_u%d" % t)(_ptr(buf), _ptr(out), bw, int(base) & ((1 << t) - 1))
        return out

    def falp(self, packed, bw, base, f, e, value_bytes=8):
        ft, ut, it, sfx = _abi.value_types(value_bytes)
        buf = np.zeros(1024, dtype=ut)
        buf[: packed.shape[0]] = packed
        out = np.zeros(1024, dtype=ft)
        self._fn("falp_" + sfx)(_ptr(buf), _ptr(out), bw, int(base) & ((1 << (8 * value_bytes)) - 1), f, e)
        return out

    def decode(self, enc, f, e):
        enc = np.ascontiguousarray(enc)
        ft, ut, it, sfx = _abi.value_types(enc.dtype.itemsize)
        out = np.zeros(1024, dtype=ft)
        self._fn("decode_" + sfx)(_ptr(enc), f, e, _ptr(out))
        return out

    def patch(self, out, exc, pos):
        sfx = _abi.value_types(out.dtype.itemsize)[3]
        exc = np.ascontiguousarray(exc)
        pos = np.ascontiguousarray(pos, dtype=np.uint16)
        self._fn("patch_" + sfx)(_ptr(out), _ptr(exc), _ptr(pos), len(pos))
        return out

    # -- ALP_RD vector primitives -----------------------------------------------------------------
    def rd_encode(self, vec, state):
        vec = np.ascontiguousarray(vec)
        ft, ut, it, sfx = _abi.value_types(vec.dtype.itemsize)
        exc = np.zeros(1024, dtype=np.uint16)
        pos = np.zeros(1024, dtype=np.uint16)
        cnt = np.zeros(1, dtype=np.uint16)
        right = np.zeros(1024, dtype=ut)
        left = np.zeros(1024, dtype=np.uint16)
        self._fn("rd_encode_" + sfx)(_ptr(vec), _ptr(state), _ptr(exc), _ptr(pos), _ptr(cnt), _ptr(right), _ptr(left))
        n = int(cnt[0])
        return dict(right=right, left=left, exc=exc[:n].copy(), pos=pos[:n].copy(), cnt=n)

    def rd_decode(self, right, left, exc, pos, state):
        right = np.ascontiguousarray(right)
        ft, ut, it, sfx = _abi.value_types(right.dtype.itemsize)
        out = np.zeros(1024, dtype=ft)
        exc = np.ascontiguousarray(exc, dtype=np.uint16)
        pos = np.ascontiguousarray(pos, dtype=np.uint16)
        left = np.ascontiguousarray(left, dtype=np.uint16)
        self._fn("rd_decode_" + sfx)(_ptr(out), _ptr(right), _ptr(left), _ptr(exc), _ptr(pos), len(pos), _ptr(state))
        return out

    # -- whole columns --------------------------------------------------------------------------
    def encode_column(self, values, n_threads=1, packed_capacity=None, exc_capacity=None):
        """The caller's row-group loop (benchmarks/benchmark.cpp:200-285) → HostColumn."""
        values = np.ascontiguousarray(values)
        sfx = _abi.value_types(values.dtype.itemsize)[3]
        n_vec = values.shape[0] // 1024
        col = _abi.HostColumn(n_vec, values.dtype.itemsize, packed_capacity, exc_capacity)
        st = col.as_struct()
        rc = self._fn("encode_column_" + sfx)(_ptr(values), values.shape[0], n_threads, ctypes.byref(st))
        if rc != 0:
            raise RuntimeError("encode_column failed: %d" % rc)
        return col

    def decode_column(self, col, first=0, n=None, n_threads=1, out=None):
        ft, ut, it, sfx = _abi.value_types(col.value_bytes)
        n = col.n_vectors - first if n is None else n
        if out is None:
            out = np.empty(n * 1024, dtype=ft)
        st = col.as_struct()
        rc = self._fn("decode_column_" + sfx)(ctypes.byref(st), first, n, n_threads, _ptr(out))
        if rc != 0:
            raise RuntimeError("decode_column failed: %d" % rc)
        return out


    # -- bench drivers (bench.py's CPU legs) ------------------------------------------------------
    def bench_init(self, values, n_threads=1, states=None):
        """encoder<PT>::init (+ rd_encoder<PT>::init) for every row-group → RG_STATE array."""
        values = np.ascontiguousarray(values)
        sfx = _abi.value_types(values.dtype.itemsize)[3]
        n_rg = -(-(values.shape[0] // 1024) // _abi.ROWGROUP_VECTORS)
        if states is None:
            states = np.zeros(max(n_rg, 1), dtype=_abi.RG_STATE_DTYPE)
        rc = self._fn("bench_encode_" + sfx)(_ptr(values), values.shape[0], n_threads, None, _ptr(states), 1)
        if rc != 0:
            raise RuntimeError("bench_init failed: %d" % rc)
        return states

    def bench_encode(self, values, n_threads=1, states=None, col=None):
        """Multi-threaded encode for timing: every thread writes into its own slice of `col` (records carry absolute
        offsets; slices leave gaps).  states=None runs the row-group init inside the call (init + encode)."""
        values = np.ascontiguousarray(values)
        vb = values.dtype.itemsize
        sfx = _abi.value_types(vb)[3]
        n_vec = values.shape[0] // 1024
        if col is None:
            units = 67 if vb == 8 else 35
            col = _abi.HostColumn(n_vec, vb, n_vec * units * 128 + n_threads * (units * 128 + 128), n_vec * 1024 + n_threads * 1024)
        st = col.as_struct()
        mode = 2 | (1 if states is None else 0)
        rc = self._fn("bench_encode_" + sfx)(_ptr(values), values.shape[0], n_threads, ctypes.byref(st), None if states is None else _ptr(states), mode)
        if rc != 0:
            raise RuntimeError("bench_encode failed: %d" % rc)
        return col

    def sum_column(self, col, first=0, n=None, n_threads=1):
        """The reference's scan query: alp_func (falp + patch_exceptions) + aggr_plus per vector, n_threads workers."""
        sfx = _abi.value_types(col.value_bytes)[3]
        n = col.n_vectors - first if n is None else n
        out = ctypes.c_double(0.0)
        st = col.as_struct()
        rc = self._fn("sum_column_" + sfx)(ctypes.byref(st), first, n, n_threads, ctypes.byref(out))
        if rc != 0:
            raise RuntimeError("sum_column failed: %d" % rc)
        return float(out.value)


_PORT = None
_REF = "unset"


def port():
    global _PORT
    if _PORT is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _PORT = CpuCodec(path, "alpo_", "port")
        f = _PORT.lib.alpo_generate_f64
        f.argtypes = [_P, _c.c_uint64, _c.c_uint64, _c.c_uint64, _c.c_int]
        f = _PORT.lib.alpo_generate_f32
        f.argtypes = [_P, _c.c_uint64, _c.c_uint64, _c.c_uint64, _c.c_int]
    return _PORT


def reference():
    """The compiled reference, or None when oracle/_ref was never built (and /root/reference is absent)."""
    global _REF
    if _REF == "unset":
        names = ["libalp_ref_v4.so", "libalp_ref_v3.so"] if _cpu_has_avx512() else ["libalp_ref_v3.so"]
        _REF = None
        for name in names:
            path = os.path.join(_HERE, "_ref", name)
            if not os.path.exists(path) and os.path.exists("/root/reference/include/alp.hpp"):
                build()
            if os.path.exists(path):
                _REF = CpuCodec(path, "alpref_", "reference")
                break
    return _REF


def best():
    """The strongest checker available: the compiled reference, else the port."""
    return reference() or port()


def generate(n_values, kind, seed=None, first_index=0, n_threads=1):
    """CPU twin of alpb200_generate_* (SURVEY.md §8d): kind 2/3 → f64, kind 4 → f32.  The generator is stateless per
    index, so n_threads > 1 simply fills disjoint slices concurrently (ctypes releases the GIL)."""
    seeds = {2: 42, 3: 43, 4: 44}
    seed = seeds[kind] if seed is None else seed
    lib = port().lib
    out = np.empty(n_values, dtype=np.float32 if kind == 4 else np.float64)
    fn = lib.alpo_generate_f32 if kind == 4 else lib.alpo_generate_f64

    def fill(lo, hi):
        if hi > lo:
            fn(_ptr(out[lo:hi]), hi - lo, first_index + lo, seed, kind)

    if n_threads <= 1 or n_values < (1 << 20):
        fill(0, n_values)
    else:
        from concurrent.futures import ThreadPoolExecutor

        step = -(-n_values // n_threads)
        with ThreadPoolExecutor(max_workers=n_threads) as pool:
            list(pool.map(lambda k: fill(k * step, min(n_values, (k + 1) * step)), range(n_threads)))
    return out
