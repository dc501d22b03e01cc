"""include/alp_b200.hpp — the reference's C++ primitive API over the C ABI.

tests/cpp/shim_roundtrip.cpp is written against the reference's own call sequence (test/test_alp_sample.cpp:137-179)
and nothing else; here it is compiled against the shim header, linked to libalp_b200.so and — on the GPU box — run
over every golden fixture vector, checking the round trip and the reference test's two golden asserts.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = os.path.join(str(tmp_path), "shim_roundtrip")
    lib_dir = os.path.join(ROOT, "alp_b200")
    cmd = [
        "g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
        os.path.join(ROOT, "tests", "cpp", "shim_roundtrip.cpp"), "-o", exe,
        "-L" + lib_dir, "-lalp_b200", "-Wl,-rpath," + lib_dir,
    ]  # fmt: skip
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def _write_cases(path, golden_vectors):
    g = golden_vectors
    with open(path, "wb") as fh:
        for c in g.index:
            x = g[c["id"] + "_input"]
            is_float = x.dtype == np.float32
            fh.write(struct.pack("<4I", int(is_float), c["golden_bw"], c["golden_exceptions"], c["scheme"]))
            raw = x.tobytes()
            fh.write(raw + b"\0" * (8192 - len(raw)))
    return len(g.index)


def test_reference_style_program_compiles_against_the_shim(tmp_path):
    _build(tmp_path)


def test_shim_fails_loudly_without_gpu(tmp_path, golden_vectors):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    exe = _build(tmp_path)
    cases = os.path.join(str(tmp_path), "cases.bin")
    _write_cases(cases, golden_vectors)
    res = subprocess.run([exe, cases], capture_output=True, text=True)
    assert res.returncode == 3 and "gpu_error" in res.stdout  # alp::gpu_error, not a silent CPU path


@pytest.mark.gpu
def test_reference_style_program_round_trips_every_fixture(tmp_path, golden_vectors):
    exe = _build(tmp_path)
    cases = os.path.join(str(tmp_path), "cases.bin")
    n = _write_cases(cases, golden_vectors)
    res = subprocess.run([exe, cases], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "cases %d bad 0" % n in res.stdout
