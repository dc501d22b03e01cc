"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/alp_b200.h declares,
agrees with the Python images of its structs, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "alp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(alpb200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import alp_b200
    from alp_b200 import _lib

    names = _declared_functions()
    assert len(names) >= 40
    for name in names:
        assert hasattr(alp_b200.lib, name), "libalp_b200.so does not export %s" % name
    # and the ctypes table covers the header exactly (no stale or missing prototypes)
    assert sorted(_lib.SIGNATURES) == names


def test_struct_layouts_match():
    import alp_b200
    from alp_b200 import _abi

    sizes = (ctypes.c_uint32 * 3)()
    alp_b200.lib.alpb200_abi_sizes(sizes)
    assert tuple(sizes) == (_abi.RG_STATE_DTYPE.itemsize, _abi.VEC_META_DTYPE.itemsize, ctypes.sizeof(_abi.Column))
    assert _abi.RG_STATE_DTYPE.fields["extra_key"][1] == 44 and _abi.RG_STATE_DTYPE.fields["dict"][1] == 24
    assert _abi.VEC_META_DTYPE.fields["packed_off"][1] == 16 and _abi.VEC_META_DTYPE.fields["bw"][1] == 27


def test_workspace_queries_need_no_gpu():
    import alp_b200

    assert alp_b200.lib.alpb200_version() >= 100
    assert alp_b200.lib.alpb200_encode_workspace_bytes(1 << 20) >= (1 << 17) * 8
    assert alp_b200.lib.alpb200_init_workspace_bytes(1 << 30) >= 10486 * 9 * 16


def test_no_cpu_fallback_without_a_gpu():
    import torch

    import alp_b200

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(alp_b200.AlpError) as err:
        alp_b200.device_count()
    assert err.value.code == alp_b200._abi.ENODEVICE
    # a compute entry point must fail too, not quietly compute on the host
    x = np.arange(1024, dtype=np.float64)
    with pytest.raises(alp_b200.AlpError):
        alp_b200.primitives.analyze_ffor(x.view(np.int64))
    with pytest.raises(alp_b200.AlpError):
        alp_b200.HostCodec(16, 8)


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under alp_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "alp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                for needle in ("pyoracle", "liboracle", "alpo_", "alpref_", "libalp_ref", "import oracle", "from oracle"):
                    if needle in text:
                        # the device headers mention the oracle only in comments that explain a constant
                        lines = [ln for ln in text.splitlines() if needle in ln and not ln.strip().startswith(("//", "*", "#", '"'))]
                        assert not lines, (f, needle, lines)
