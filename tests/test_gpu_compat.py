"""Runs the C++ program that mirrors the reference's tests/test_fhe.cu against the compat headers (include/fhe/*.cuh)
on the GPU.  The binary is built by __graft_entry__.build() (nvcc, in-tree)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_fhe_compat.bin")


def test_reference_test_program_against_compat_headers():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    assert os.path.exists(BIN), "tests/cpp/test_fhe_compat.bin missing: run __graft_entry__.build()"
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL COMPAT TESTS PASSED" in r.stdout
    for part in ("bigint arithmetic ok", "ntt round trip ok", "polynomial multiplication ok", "rns context ok",
                 "fhe operations ok", "error behaviour ok"):
        assert part in r.stdout
