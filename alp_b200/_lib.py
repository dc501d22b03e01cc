"""Loads libalp_b200.so (built in-tree by __graft_entry__.build()) and declares the C ABI to ctypes."""
import ctypes
import os

from . import _abi

# ALPB200_LIB lets a developer load an experimental build of the same library (tools/decode_probe.py)
LIB_PATH = os.environ.get("ALPB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libalp_b200.so")

_c = ctypes
_P = ctypes.c_void_p


class AlpError(RuntimeError):
    """A C-ABI call returned a negative ALPB200_E* code."""

    def __init__(self, code, message):
        super().__init__("alp_b200 error %d: %s" % (code, message))
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "alp_b200: %s is missing. The CUDA library is the product and there is no CPU fallback; build it with "
        "`python -c 'import __graft_entry__ as g; g.build()'` from the repo root." % LIB_PATH
    )

lib = ctypes.CDLL(LIB_PATH)

# every exported symbol of include/alp_b200.h with its signature; tests/test_abi.py checks the list against the header
SIGNATURES = {
    "alpb200_version": ([], _c.c_int),
    "alpb200_abi_sizes": ([_P], None),
    "alpb200_last_error": ([], _c.c_char_p),
    "alpb200_device_count": ([], _c.c_int),
    "alpb200_init_workspace_bytes": ([_c.c_uint64], _c.c_size_t),
    "alpb200_rowgroup_init_f64": ([_P, _c.c_uint64, _P, _P, _P], _c.c_int),
    "alpb200_rowgroup_init_f32": ([_P, _c.c_uint64, _P, _P, _P], _c.c_int),
    "alpb200_encode_workspace_bytes": ([_c.c_uint64], _c.c_size_t),
    "alpb200_encode_f64": ([_P, _c.c_uint64, _P, _P, _P, _P], _c.c_int),
    "alpb200_encode_f32": ([_P, _c.c_uint64, _P, _P, _P, _P], _c.c_int),
    "alpb200_encode_ex_f64": ([_P, _c.c_uint64, _P, _P, _P, _P, _c.c_uint32], _c.c_int),
    "alpb200_encode_ex_f32": ([_P, _c.c_uint64, _P, _P, _P, _P, _c.c_uint32], _c.c_int),
    "alpb200_encode_unordered_f64": ([_P, _c.c_uint64, _P, _P, _P, _P], _c.c_int),
    "alpb200_encode_unordered_f32": ([_P, _c.c_uint64, _P, _P, _P, _P], _c.c_int),
    "alpb200_decode_f64": ([_P, _c.c_uint64, _c.c_uint64, _P, _P], _c.c_int),
    "alpb200_decode_f32": ([_P, _c.c_uint64, _c.c_uint64, _P, _P], _c.c_int),
    "alpb200_decode_values_f64": ([_P, _c.c_uint64, _c.c_uint64, _P, _P, _P], _c.c_int),
    "alpb200_decode_values_f32": ([_P, _c.c_uint64, _c.c_uint64, _P, _P, _P], _c.c_int),
    "alpb200_fill_invalid_f64": ([_P, _c.c_uint64, _P, _P, _P], _c.c_int),
    "alpb200_fill_invalid_f32": ([_P, _c.c_uint64, _P, _P, _P], _c.c_int),
    "alpb200_decode_sum_f64": ([_P, _c.c_uint64, _c.c_uint64, _P, _P], _c.c_int),
    "alpb200_decode_sum_f32": ([_P, _c.c_uint64, _c.c_uint64, _P, _P], _c.c_int),
    "alpb200_decode_minmax_f64": ([_P, _c.c_uint64, _c.c_uint64, _P, _P], _c.c_int),
    "alpb200_decode_minmax_f32": ([_P, _c.c_uint64, _c.c_uint64, _P, _P], _c.c_int),
    "alpb200_decode_filter_f64": ([_P, _c.c_uint64, _c.c_uint64, _c.c_uint32, _c.c_double, _P, _P, _P], _c.c_int),
    "alpb200_decode_filter_f32": ([_P, _c.c_uint64, _c.c_uint64, _c.c_uint32, _c.c_double, _P, _P, _P], _c.c_int),
    "alpb200_decode_sum_ex_f64": ([_P, _c.c_uint64, _c.c_uint64, _P, _c.c_uint32, _P], _c.c_int),
    "alpb200_decode_sum_ex_f32": ([_P, _c.c_uint64, _c.c_uint64, _P, _c.c_uint32, _P], _c.c_int),
    "alpb200_ctx_create": ([_P, _c.c_int, _c.c_uint64, _c.c_int], _c.c_int),
    "alpb200_ctx_create_ex": ([_P, _c.c_int, _c.c_uint64, _c.c_int, _c.c_uint64, _c.c_uint64], _c.c_int),
    "alpb200_column_validate_host": ([_P, _c.c_int], _c.c_int),
    "alpb200_column_validate_device": ([_P, _c.c_int, _P, _P], _c.c_int),
    "alpb200_ctx_destroy": ([_P], None),
    "alpb200_ctx_set_option": ([_P, _c.c_int, _c.c_int], _c.c_int),
    "alpb200_compress_host_f64": ([_P, _P, _c.c_uint64, _P], _c.c_int),
    "alpb200_compress_host_f32": ([_P, _P, _c.c_uint64, _P], _c.c_int),
    "alpb200_decompress_host_f64": ([_P, _P, _P], _c.c_int),
    "alpb200_decompress_host_f32": ([_P, _P, _P], _c.c_int),
    "alpb200_sum_host_f64": ([_P, _P, _P], _c.c_int),
    "alpb200_sum_host_f32": ([_P, _P, _P], _c.c_int),
    "alpb200_host_alloc": ([_c.c_size_t], _P),
    "alpb200_host_free": ([_P], None),
    "alpb200_prim_encode_f64": ([_P] * 8, _c.c_int),
    "alpb200_prim_encode_f32": ([_P] * 8, _c.c_int),
    "alpb200_prim_analyze_ffor_i64": ([_P, _P, _P], _c.c_int),
    "alpb200_prim_analyze_ffor_i32": ([_P, _P, _P], _c.c_int),
    "alpb200_prim_ffor_u64": ([_P, _P, _c.c_uint8, _c.c_uint64], _c.c_int),
    "alpb200_prim_ffor_u32": ([_P, _P, _c.c_uint8, _c.c_uint32], _c.c_int),
    "alpb200_prim_ffor_u16": ([_P, _P, _c.c_uint8, _c.c_uint16], _c.c_int),
    "alpb200_prim_ffor_u8": ([_P, _P, _c.c_uint8, _c.c_uint8], _c.c_int),
    "alpb200_prim_unffor_u64": ([_P, _P, _c.c_uint8, _c.c_uint64], _c.c_int),
    "alpb200_prim_unffor_u32": ([_P, _P, _c.c_uint8, _c.c_uint32], _c.c_int),
    "alpb200_prim_unffor_u16": ([_P, _P, _c.c_uint8, _c.c_uint16], _c.c_int),
    "alpb200_prim_unffor_u8": ([_P, _P, _c.c_uint8, _c.c_uint8], _c.c_int),
    "alpb200_prim_falp_f64": ([_P, _P, _c.c_uint8, _c.c_uint64, _c.c_uint8, _c.c_uint8], _c.c_int),
    "alpb200_prim_falp_f32": ([_P, _P, _c.c_uint8, _c.c_uint32, _c.c_uint8, _c.c_uint8], _c.c_int),
    "alpb200_prim_decode_f64": ([_P, _c.c_uint8, _c.c_uint8, _P], _c.c_int),
    "alpb200_prim_decode_f32": ([_P, _c.c_uint8, _c.c_uint8, _P], _c.c_int),
    "alpb200_prim_patch_f64": ([_P, _P, _P, _c.c_uint16], _c.c_int),
    "alpb200_prim_patch_f32": ([_P, _P, _P, _c.c_uint16], _c.c_int),
    "alpb200_prim_rd_encode_f64": ([_P] * 7, _c.c_int),
    "alpb200_prim_rd_encode_f32": ([_P] * 7, _c.c_int),
    "alpb200_prim_rd_decode_f64": ([_P, _P, _P, _P, _P, _c.c_uint16, _P], _c.c_int),
    "alpb200_prim_rd_decode_f32": ([_P, _P, _P, _P, _P, _c.c_uint16, _P], _c.c_int),
    "alpb200_prim_init_f64": ([_P, _c.c_uint64, _c.c_uint64, _P], _c.c_int),
    "alpb200_prim_init_f32": ([_P, _c.c_uint64, _c.c_uint64, _P], _c.c_int),
    "alpb200_prim_rd_init_f64": ([_P, _c.c_uint64, _c.c_uint64, _P], _c.c_int),
    "alpb200_prim_rd_init_f32": ([_P, _c.c_uint64, _c.c_uint64, _P], _c.c_int),
    "alpb200_generate_f64": ([_P, _c.c_uint64, _c.c_uint64, _c.c_uint64, _c.c_int, _P], _c.c_int),
    "alpb200_generate_f32": ([_P, _c.c_uint64, _c.c_uint64, _c.c_uint64, _c.c_int, _P], _c.c_int),
}

for _name, (_args, _res) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = the library does not export what the header declares
    _fn.argtypes = _args
    _fn.restype = _res

_sizes = (ctypes.c_uint32 * 3)()
lib.alpb200_abi_sizes(_sizes)
if tuple(_sizes) != (_abi.RG_STATE_DTYPE.itemsize, _abi.VEC_META_DTYPE.itemsize, ctypes.sizeof(_abi.Column)):
    raise ImportError("alp_b200: %s was built against a different include/alp_b200.h; rebuild it" % LIB_PATH)


def check(rc):
    """Raise AlpError for a negative return code."""
    if rc < 0:
        raise AlpError(rc, lib.alpb200_last_error().decode(errors="replace"))
    return rc
