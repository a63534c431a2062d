"""Parity against the REFERENCE ITSELF, for the part of it that is complete: its add_mod / sub_mod / mul_mod_montgomery device
functions (include/bigint.cuh:27-140) run through its own batch kernels (src/bigint.cu:171-215), compiled from the reference
sources into oracle/_ref/libref_kernels.so by oracle/Makefile (target _ref) in the build container and shipped prebuilt to the GPU
box.  Both the CPU oracle's modular arithmetic and the CUDA element-wise ops behind the C ABI are compared with them bit for bit.
The reference never enters or leaves the Montgomery domain: its product kernel returns a*b*2^-256 mod q, so the
comparison multiplies by 2^256 mod q (exact Python integers)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
R256 = 1 << 256


@pytest.fixture(scope="module")
def ref(oracle):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if oracle.ref_kernels() is None:
        pytest.skip("oracle/_ref/libref_kernels.so was not built (needs /root/reference at build time)")
    return oracle


def _operands(rng, q, n=4096):
    a = rng.integers(0, q, n, dtype=np.uint64); b = rng.integers(0, q, n, dtype=np.uint64)
    a[:4] = [0, q - 1, q - 1, 1]; b[:4] = [0, q - 1, 1, q - 1]          # edge vectors
    return a, b


def test_reference_known_answers(ref):
    """tests/test_fhe.cu:34-56 prints these without asserting them"""
    a = np.array([12345], dtype=np.uint64); b = np.array([67890], dtype=np.uint64)
    assert int(ref.ref_batch("add", a, b, 100000)[0][0]) == 80235
    assert int(ref.ref_batch("sub", a, b, 100000)[0][0]) == 44455


@pytest.mark.parametrize("q", [100000, 12289, 40961, "chain0", "chain7", "q61"])
def test_oracle_arithmetic_equals_reference_kernels(ref, oracle, q):
    if q == "q61":
        q = (1 << 61) - (1 << 18) * 3 + 1            # any odd 61-bit modulus: the kernels do not need a prime
    elif isinstance(q, str):
        q = oracle.prime_chain(8)[int(q[5:])]
    rng = np.random.default_rng(q % 1000003)
    a, b = _operands(rng, q)
    out, fits = ref.ref_batch("add", a, b, q)
    assert fits and [int(v) for v in out] == [oracle.add_mod(int(x), int(y), q) for x, y in zip(a, b)]
    out, fits = ref.ref_batch("sub", a, b, q)
    assert fits and [int(v) for v in out] == [oracle.sub_mod(int(x), int(y), q) for x, y in zip(a, b)]
    if q % 2:
        out, fits = ref.ref_batch("mul_montgomery", a, b, q)
        r = R256 % q
        assert fits and [int(v) * r % q for v in out] == [oracle.mul_mod(int(x), int(y), q) for x, y in zip(a, b)]


def test_cuda_elementwise_ops_equal_reference_kernels(ref, oracle):
    """fhe_b200_poly_add / _sub / _mul through the C ABI vs the reference's kernels on the same operands, limb by limb"""
    import fhe_b200
    from fhe_b200.engine import to_device, to_host
    n, chain = 4096, oracle.prime_chain(3)
    plan = fhe_b200.Plan(n, chain)
    rng = np.random.default_rng(99)
    ops = [_operands(rng, q, n) for q in chain]
    a = np.stack([o[0] for o in ops])[None]; b = np.stack([o[1] for o in ops])[None]
    da, db = to_device(a), to_device(b)
    got = {"add": to_host(plan.add(da, db))[0], "sub": to_host(plan.sub(da, db))[0], "mul": to_host(plan.mul(da, db))[0]}
    for i, q in enumerate(chain):
        assert np.array_equal(got["add"][i], ref.ref_batch("add", a[0, i], b[0, i], q)[0]), ("add", i)
        assert np.array_equal(got["sub"][i], ref.ref_batch("sub", a[0, i], b[0, i], q)[0]), ("sub", i)
        mont, fits = ref.ref_batch("mul_montgomery", a[0, i], b[0, i], q)
        r = R256 % q
        assert fits and [int(v) for v in got["mul"][i]] == [int(v) * r % q for v in mont], ("mul", i)


@pytest.mark.parametrize("n,q", [(1024, 12289), (2048, 40961), (4096, "chain0")])
def test_transform_built_from_reference_arithmetic_equals_cuda_ntt(ref, oracle, n, q):
    """The negacyclic transform evaluated stage by stage with EVERY modular operation done by the reference's own kernels
    (products through its Montgomery kernel with the twiddle pre-multiplied by 2^256, sums and differences through its add/sub
    kernels) equals the CUDA NTT behind the C ABI, and the oracle, word for word.  The reference's test sizes (tests/test_fhe.cu:
    68-116, 126-167) and one 60-bit prime.  Twiddles and ordering are the engine's (the reference's own tables are placeholders)."""
    import fhe_b200
    from fhe_b200.engine import to_device, to_host
    if isinstance(q, str):
        q = oracle.prime_chain(1)[0]
    lg = n.bit_length() - 1
    psi = oracle.find_psi(q, n)
    brev = lambda x: int(format(x, f"0{lg}b")[::-1], 2)
    r = R256 % q
    psi_rev_r = np.array([pow(psi, brev(i), q) * r % q for i in range(n)], dtype=np.uint64)     # w * 2^256 mod q
    rng = np.random.default_rng(n)
    a = rng.integers(0, q, n, dtype=np.uint64)
    a[:3] = [0, q - 1, 1]
    x = a.copy()
    t, m = n, 1
    while m < n:
        t >>= 1
        lo = (np.arange(m)[:, None] * 2 * t + np.arange(t)[None, :]).reshape(-1)
        w = np.repeat(psi_rev_r[m:2 * m], t)
        v, fits = ref.ref_batch("mul_montgomery", x[lo + t], w, q)
        assert fits
        u = x[lo].copy()
        x[lo] = ref.ref_batch("add", u, v, q)[0]
        x[lo + t] = ref.ref_batch("sub", u, v, q)[0]
        m <<= 1
    assert np.array_equal(x, oracle.ntt_forward(a, q))
    plan = fhe_b200.Plan(n, [q])
    d = to_device(a[None, None])
    plan.forward(d)
    torch.cuda.synchronize()
    assert np.array_equal(to_host(d)[0, 0], x)
