"""Pin the CPU oracle (NTT / polynomial part) against the reference's own known answers
(tests/test_fhe.cu:34-56, 68-116, 126-167) and against two independent
definitions (schoolbook product, definitional transform)."""
import numpy as np
import pytest


def test_reference_bigint_known_answers(oracle):
    # /root/reference/tests/test_fhe.cu:34-56 : 12345 +- 67890 mod 100000 (true values 80235, 44455)
    assert oracle.add_mod(12345, 67890, 100000) == 80235
    assert oracle.sub_mod(12345, 67890, 100000) == 44455


def test_prime_chain_matches_survey(oracle, chain):
    # values computed independently during the survey
    assert chain[0] == 1152921504606584833 == 0xFFFFFFFFFFC0001
    assert chain[1] == 1152921504598720513
    assert chain[2] == 1152921504592429057
    assert chain[3] == 1152921504581419009
    assert chain[31] == 1152921504455589889
    assert chain[39] == 1152921504396869633
    for p in chain:
        assert p < 2**60 and p % 2**18 == 1 and oracle.is_prime(p)
        assert pow(2, p - 1, p) == 1
    assert sorted(chain, reverse=True) == chain and len(set(chain)) == len(chain)


@pytest.mark.parametrize("n,q", [(1024, 12289), (2048, 40961), (4096, 0xFFFFFFFFFFC0001)])
def test_psi_rule(oracle, n, q):
    psi = oracle.find_psi(q, n)
    assert pow(psi, n, q) == q - 1          # order exactly 2N
    assert pow(psi, 2 * n, q) == 1
    x = next(x for x in range(2, 1000) if pow(x, (q - 1) // 2, q) == q - 1)
    assert psi == pow(x, (q - 1) // (2 * n), q)


@pytest.mark.parametrize("n,q", [(8, 17), (16, 97), (64, 12289), (256, 12289), (1024, 12289)])
def test_forward_matches_definition_and_ordering(oracle, n, q):
    rng = np.random.default_rng(1)
    a = rng.integers(0, q, n, dtype=np.uint64)
    X = oracle.negacyclic_dft_def(a, q)           # natural order
    # python restatement of the definition as a third opinion at tiny n
    if n <= 64:
        psi = oracle.find_psi(q, n)
        Xp = [sum(int(a[j]) * pow(psi, j * (2 * k + 1), q) for j in range(n)) % q for k in range(n)]
        assert [int(v) for v in X] == Xp
    Y = oracle.ntt_forward(a, q)                  # bit-reversed order
    br = oracle.bitrev_perm(n)
    assert np.array_equal(Y, X[br])
    assert np.array_equal(oracle.ntt_inverse(Y, q), a)


def test_reference_roundtrip_vector(oracle):
    # /root/reference/tests/test_fhe.cu:68-116 : N=1024, q=12289, x_i = i+1, inverse(forward(x)) == x
    n, q = 1024, 12289
    x = np.arange(1, n + 1, dtype=np.uint64)
    assert np.array_equal(oracle.ntt_inverse(oracle.ntt_forward(x, q), q), x)


def test_reference_polymul_config(oracle):
    # /root/reference/tests/test_fhe.cu:126-167 : N=2048, q=40961, coefficients rand()%100 (product never checked there)
    n, q = 2048, 40961
    rng = np.random.default_rng(2)
    a = rng.integers(0, 100, n, dtype=np.uint64); b = rng.integers(0, 100, n, dtype=np.uint64)
    assert np.array_equal(oracle.negacyclic_mul_ntt(a, b, q), oracle.schoolbook_negacyclic(a, b, q))


def test_schoolbook_against_python(oracle):
    n, q = 32, 0xFFFFFFFFFFC0001
    rng = np.random.default_rng(3)
    a = rng.integers(0, q, n, dtype=np.uint64); b = rng.integers(0, q, n, dtype=np.uint64)
    ref = [0] * n
    for i in range(n):
        for j in range(n):
            k = i + j
            v = int(a[i]) * int(b[j])
            if k >= n:
                ref[k - n] -= v
            else:
                ref[k] += v
    ref = [v % q for v in ref]
    assert [int(v) for v in oracle.schoolbook_negacyclic(a, b, q)] == ref


def test_config1_ntt_mul_equals_schoolbook(oracle, chain):
    # BASELINE.json config 1: N=4096, one 60-bit prime, fwd -> pointwise -> inv == schoolbook, bit exact
    n, q = 4096, chain[0]
    rng = np.random.default_rng(0x5EED0001)
    a = rng.integers(0, q, n, dtype=np.uint64); b = rng.integers(0, q, n, dtype=np.uint64)
    assert np.array_equal(oracle.negacyclic_mul_ntt(a, b, q), oracle.schoolbook_negacyclic(a, b, q))


@pytest.mark.parametrize("vec", ["zero", "qm1", "delta0", "deltaN"])
def test_edge_vectors(oracle, chain, vec):
    n, q = 1024, chain[5]
    a = np.zeros(n, dtype=np.uint64)
    if vec == "qm1": a[:] = q - 1
    if vec == "delta0": a[0] = 1
    if vec == "deltaN": a[n - 1] = 1
    Y = oracle.ntt_forward(a, q)
    if vec == "zero": assert not Y.any()
    if vec == "delta0": assert (Y == 1).all()
    assert np.array_equal(oracle.ntt_inverse(Y, q), a)
    assert np.array_equal(Y, oracle.negacyclic_dft_def(a, q)[oracle.bitrev_perm(n)])


def test_batched_shoup_baseline_equals_plain(oracle, chain):
    n, limbs, batch = 2048, 3, 4
    mods = chain[:limbs]
    rng = np.random.default_rng(4)
    x = np.stack([np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in mods]) for _ in range(batch)])
    eng = oracle.RnsNtt(n, mods)
    y = eng.forward(x, threads=2)
    for b in range(batch):
        for l, q in enumerate(mods):
            assert np.array_equal(y[b, l], oracle.ntt_forward(x[b, l], q))
    assert np.array_equal(eng.inverse(y, threads=2), x)


def test_elementwise(oracle, chain):
    n, mods = 64, chain[:2]
    rng = np.random.default_rng(5)
    a = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in mods])[None]
    b = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in mods])[None]
    for l, q in enumerate(mods):
        A = [int(v) for v in a[0, l]]; B = [int(v) for v in b[0, l]]
        assert [int(v) for v in oracle.poly_add(a, b, mods, n)[0, l]] == [(x + y) % q for x, y in zip(A, B)]
        assert [int(v) for v in oracle.poly_sub(a, b, mods, n)[0, l]] == [(x - y) % q for x, y in zip(A, B)]
        assert [int(v) for v in oracle.poly_mul(a, b, mods, n)[0, l]] == [(x * y) % q for x, y in zip(A, B)]
        assert [int(v) for v in oracle.poly_mac(a, a, b, mods, n)[0, l]] == [(x + x * y) % q for x, y in zip(A, B)]
        assert [int(v) for v in oracle.poly_mul_scalar(a, [7, 9], mods, n)[0, l]] == [(x * [7, 9][l]) % q for x in A]
