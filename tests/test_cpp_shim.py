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


# ---- the reference's OWN unit test, unchanged, against the shim -----------------------------------------------------
REF_TEST = os.path.join(ROOT, "oracle", "_ref", "ref_test_on_shim")
FIXTURE_ROOT = "/tmp/alp_b200_ref_fixtures"  # baked into the binary as ALP_CMAKE_SOURCE_DIR (oracle/Makefile)


def _materialise_reference_fixtures(golden_vectors):
    """Recreate data/<dir>/<file>.csv of the reference tree from the golden inputs (first 1024 values of each column,
    all the reference's test reads), written so that std::stod / std::stof give back the exact values."""
    g = golden_vectors
    done = {}
    for c in sorted(g.index, key=lambda c: c["dtype"] != "float64"):  # a file shared by a double and a float case: the double text
        path = os.path.join(FIXTURE_ROOT, c["relpath"])
        if path in done:
            continue
        os.makedirs(os.path.dirname(path), exist_ok=True)
        x = g[c["id"] + "_input"]
        with open(path, "w") as fh:
            for v in x:
                fh.write((repr(float(v)) if x.dtype == np.float64 else str(v)) + "\n")
        done[path] = c["id"]
    return len(done)


@pytest.mark.gpu
def test_reference_own_unit_test_passes_on_the_shim(golden_vectors):
    """test/test_alp_sample.cpp of the reference — compiled where it lies, not a line changed — linked against
    include/alp_b200.hpp + libalp_b200.so: its six gtest cases (round trip of 103 columns + the golden bit_width /
    exceptions_count asserts, :172-179) must pass with every primitive served by the GPU."""
    if not os.path.exists(REF_TEST):
        pytest.skip("oracle/_ref/ref_test_on_shim not built (needs /root/reference at build time)")
    assert _materialise_reference_fixtures(golden_vectors) >= 100
    res = subprocess.run([REF_TEST], capture_output=True, text=True, timeout=900)
    tail = res.stdout[-2000:] + res.stderr[-2000:]
    assert res.returncode == 0, tail
    assert "[  PASSED  ] 6 tests." in res.stdout, tail
    assert res.stdout.count("[       OK ]") == 6, tail


def test_reference_own_unit_test_fails_loudly_without_gpu(golden_vectors):
    """The same binary on a machine without a GPU: every case must FAIL with the shim's gpu_error — a silent CPU path
    behind the reference's API would let it pass."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not os.path.exists(REF_TEST):
        pytest.skip("oracle/_ref/ref_test_on_shim not built (needs /root/reference at build time)")
    _materialise_reference_fixtures(golden_vectors)
    res = subprocess.run([REF_TEST], capture_output=True, text=True, timeout=300)
    assert res.returncode != 0
    assert "[  PASSED  ] 0 tests." in res.stdout and res.stdout.count("[  FAILED  ]") >= 6
    assert "no CUDA device" in res.stdout or "gpu_error" in res.stdout or "CUDA" in res.stdout, res.stdout[-1500:]
