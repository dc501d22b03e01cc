"""alp_b200 — B200-native ALP / ALP_RD column codec.

Host-side mirror (Python, ctypes) of the C ABI in include/alp_b200.h, which in turn mirrors the reference's
primitive API (cwida/ALP PRIMITIVES.md).  PyTorch is used for device memory, streams and torch.distributed only;
every computation happens in the hand-written sm_100a kernels of libalp_b200.so.  There is no CPU fallback: the
package refuses to import without the CUDA library, and every call fails without a GPU.
"""
from . import _abi  # noqa: F401
from ._lib import LIB_PATH, AlpError, lib  # noqa: F401
from .codec import (  # noqa: F401
    DeviceColumn,
    HostCodec,
    decode,
    SUM_DECIMAL,
    decode_filter,
    decode_minmax,
    minmax_result,
    decode_sum,
    decode_values,
    device_count,
    encode,
    fill_invalid,
    generate,
    rowgroup_init,
)
from . import primitives  # noqa: F401

__all__ = [
    "AlpError",
    "DeviceColumn",
    "HostCodec",
    "LIB_PATH",
    "decode",
    "SUM_DECIMAL",
    "decode_filter",
    "decode_minmax",
    "minmax_result",
    "decode_sum",
    "decode_values",
    "device_count",
    "encode",
    "fill_invalid",
    "generate",
    "lib",
    "primitives",
    "rowgroup_init",
]
