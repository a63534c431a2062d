"""CPU emulation of the CUDA kernel bodies (csrc/ntt_core.cuh compiled by g++, one emulated thread at a time)
against the oracle.  This validates index maps, twiddle addressing and the lazy-reduction bounds in the GPU-less
container; the real parity tests are the -m gpu ones."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gpu-homomorphic-encryption_b200", "csrc")
EMUL = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL, "libemul.so")
    srcs = [os.path.join(EMUL, "emul_ntt.cpp"), os.path.join(CSRC, "tables.cpp"), os.path.join(CSRC, "ntt_core.cuh"),
            os.path.join(CSRC, "modarith.cuh"), os.path.join(CSRC, "ntt_bal.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        # -DFHE_CHECK_BOUNDS: every butterfly asserts its compile-time lazy bound (values < B*q, no 64-bit wrap)
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DFHE_CHECK_BOUNDS", "-I" + CSRC,
                               srcs[0], srcs[1], "-o", so])
    lib = C.CDLL(so)
    lib.emul_ntt.argtypes = [C.POINTER(C.c_uint64), C.c_uint32, C.c_uint64, C.c_int, C.c_int]
    lib.emul_ntt_bal.argtypes = lib.emul_ntt.argtypes
    lib.emul_mul_mod.restype = C.c_uint64
    lib.emul_mul_mod.argtypes = [C.c_uint64] * 3
    lib.emul_barrett128.restype = C.c_uint64
    lib.emul_barrett128.argtypes = [C.c_uint64] * 3
    return lib


def _run(lib, a, q, inverse, hb, bal=False):
    x = a.copy()
    fn = lib.emul_ntt_bal if bal else lib.emul_ntt
    assert fn(x.ctypes.data_as(C.POINTER(C.c_uint64)), x.size, q, inverse, hb) == 0
    return x


@pytest.mark.parametrize("logn", [9, 10, 11, 12, 13, 16, 17])
def test_kernel_bodies_match_oracle(emul, oracle, chain, logn):
    n = 1 << logn
    p61 = oracle.prime_chain(1, bits=61)[0]
    rng = np.random.default_rng(logn)
    for q, hb in [(chain[0], 16), (chain[7], 8), (p61, 8)]:
        for a in (rng.integers(0, q, n, dtype=np.uint64), np.full(n, q - 1, dtype=np.uint64)):
            ref = oracle.ntt_forward(a, q)
            assert np.array_equal(_run(emul, a, q, 0, hb), ref)
            assert np.array_equal(_run(emul, ref, q, 1, hb), a)


@pytest.mark.parametrize("logn", [13, 14, 15, 16])
def test_balanced_bodies_match_oracle(emul, oracle, chain, logn):
    """ntt_bal.cuh: pass A (column pass, CTA exchange) + pass B (tile pairs, warp exchange), both directions."""
    n = 1 << logn
    p61 = oracle.prime_chain(1, bits=61)[0]
    small = next(p for p in range(2 * n + 1, 1 << 30, 2 * n) if oracle.is_prime(p))       # far from 2^60: the generic csub path
    rng = np.random.default_rng(100 + logn)
    for q, hb in [(chain[0], 16), (chain[7], 8), (p61, 8), (small, 16)]:
        for a in (rng.integers(0, q, n, dtype=np.uint64), np.full(n, q - 1, dtype=np.uint64)):
            ref = oracle.ntt_forward(a, q)
            assert np.array_equal(_run(emul, a, q, 0, hb, bal=True), ref)
            assert np.array_equal(_run(emul, ref, q, 1, hb, bal=True), a)


def test_near60_boundary_prime(emul, oracle):
    """near60_reduce must hold its [0, 2q) promise for the smallest admissible modulus (2^60 - q just under 2^32), and the
    first prime below that window must take the generic path."""
    q = (1 << 60) - (1 << 32) + 1
    while not (oracle.is_prime(q) and q % (1 << 18) == 1):
        q += 1 << 18
    assert (1 << 60) - (1 << 32) < q < (1 << 60)
    for logn in (12, 16):
        n = 1 << logn
        rng = np.random.default_rng(5)
        for a in (rng.integers(0, q, n, dtype=np.uint64), np.full(n, q - 1, dtype=np.uint64)):
            ref = oracle.ntt_forward(a, q)
            assert np.array_equal(_run(emul, a, q, 0, 16), ref)
            assert np.array_equal(_run(emul, ref, q, 1, 16), a)
            if logn == 16:
                assert np.array_equal(_run(emul, a, q, 0, 16, bal=True), ref)
                assert np.array_equal(_run(emul, ref, q, 1, 16, bal=True), a)
    # just outside the window (2^60 - 2^32 - something): the emulator, like the plan, must select the generic reduction
    q2 = (1 << 60) - (1 << 32) - (1 << 18) + 1
    while not (oracle.is_prime(q2) and q2 % (1 << 18) == 1):
        q2 -= 1 << 18
    n = 1 << 13
    a = np.full(n, q2 - 1, dtype=np.uint64)
    ref = oracle.ntt_forward(a, q2)
    for bal in (False, True):
        assert np.array_equal(_run(emul, a, q2, 0, 16, bal=bal), ref)
        assert np.array_equal(_run(emul, ref, q2, 1, 16, bal=bal), a)


def test_reference_test_primes(emul, oracle):
    # the reference's own NTT test moduli: (1024, 12289) tests/test_fhe.cu:68-70, (2048, 40961) :129-131
    for n, q in [(1024, 12289), (2048, 40961)]:
        x = np.arange(1, n + 1, dtype=np.uint64) % np.uint64(q)
        y = _run(emul, x, q, 0, 16)
        assert np.array_equal(y, oracle.ntt_forward(x, q))
        assert np.array_equal(_run(emul, y, q, 1, 16), x)


def test_barrett_and_mulmod(emul, oracle, chain):
    rng = np.random.default_rng(9)
    p61 = oracle.prime_chain(1, bits=61)[0]
    for q in [chain[0], chain[31], p61, 12289, 65537, 3]:
        for _ in range(2000):
            a = int(rng.integers(0, q)); b = int(rng.integers(0, q))
            assert emul.emul_mul_mod(a, b, q) == a * b % q
            hi = int(rng.integers(0, 2**64, dtype=np.uint64)); lo = int(rng.integers(0, 2**64, dtype=np.uint64))
            assert emul.emul_barrett128(hi, lo, q) == ((hi << 64) | lo) % q
        for hi, lo in [(2**64 - 1, 2**64 - 1), (0, 0), (0, q - 1), (0, q), (q - 1, 2**64 - 1)]:
            assert emul.emul_barrett128(hi, lo, q) == ((hi << 64) | lo) % q
