"""Pin the RNS / BFV part of the CPU oracle against exact Python big-integer arithmetic."""
from math import prod

import numpy as np
import pytest

from bigint_ref import centred, crt, negacyclic_mul, poly_from_rns, poly_to_rns, round_half_up


def _rand_big(rng, lo, hi):
    span = hi - lo
    nbytes = (span.bit_length() + 7) // 8 + 8
    return lo + int.from_bytes(rng.bytes(nbytes), "little") % span


@pytest.mark.parametrize("S,T", [(1, 3), (2, 3), (4, 5), (8, 24), (24, 25)])
def test_exact_base_conversion(oracle, chain, S, T):
    src, dst = chain[:S], chain[S:S + T]
    Q = prod(src)
    lc = oracle.LinComb.conv(src, dst)
    rng = np.random.default_rng(10 + S)
    n = 64
    xs = [_rand_big(rng, -(Q // 2), Q - Q // 2 - 1) for _ in range(n - 4)] + [0, 1, -1, Q // 2 - 1]
    inp = np.array([[x % q for x in xs] for q in src], dtype=np.uint64)
    out = lc.apply(inp)
    for k, m in enumerate(dst):
        assert [int(v) for v in out[k]] == [x % m for x in xs]
    # constants against their definitions
    cst = lc.constants()
    for i, q in enumerate(src):
        assert int(cst["pre"][i]) == pow(Q // q, -1, q)
        assert (int(cst["th_hi"][i]) << 64) | int(cst["th_lo"][i]) == (1 << 128) // q
        for k, m in enumerate(dst):
            assert int(cst["M"][i, k]) == (Q // q) % m
    for k, m in enumerate(dst):
        assert int(cst["c"][k]) == (-Q) % m


@pytest.mark.parametrize("L,R,t", [(1, 2, 65537), (2, 3, 65537), (4, 5, 786433), (3, 4, 1 << 20)])
def test_scale_and_round_hps(oracle, chain, L, R, t):
    qs, ps = chain[:L], chain[L:L + R]
    Q, P = prod(qs), prod(ps)
    lc = oracle.LinComb.scale(qs, ps, t, ps, True)
    rng = np.random.default_rng(20 + L)
    n = 64
    bound = Q * P // (4 * t)   # keeps |t/Q d| < P/2
    ds = [_rand_big(rng, -bound, bound) for _ in range(n - 3)] + [0, 1, -1]
    dq = np.array([[d % q for d in ds] for q in qs], dtype=np.uint64)
    dp = np.array([[d % p for d in ds] for p in ps], dtype=np.uint64)
    out = lc.apply(dq, extra=dp)
    for k, p in enumerate(ps):
        assert [int(v) for v in out[k]] == [round_half_up(t * d, Q) % p for d in ds]


@pytest.mark.parametrize("L,t", [(1, 65537), (2, 65537), (4, 256), (24, 65537)])
def test_decrypt_scaling(oracle, chain, L, t):
    qs = chain[:L]
    Q = prod(qs)
    lc = oracle.LinComb.scale(qs, [], t, [t], False)
    rng = np.random.default_rng(30 + L)
    xs = [_rand_big(rng, 0, Q) for _ in range(64)]
    inp = np.array([[x % q for x in xs] for q in qs], dtype=np.uint64)
    out = lc.apply(inp)
    assert [int(v) for v in out[0]] == [round_half_up(t * x, Q) % t for x in xs]


def test_modswitch_drop_last(oracle, chain):
    mods = chain[:3]
    Q = prod(mods)
    rng = np.random.default_rng(40)
    xs = [_rand_big(rng, 0, Q) for _ in range(32)]
    inp = np.array([[x % q for x in xs] for q in mods], dtype=np.uint64)
    out = oracle.modswitch_drop_last(inp, mods)
    ql = mods[-1]
    for i, q in enumerate(mods[:-1]):
        # round(x / q_last) with ties (r == q_last/2 impossible: q_last odd)
        assert [int(v) for v in out[i]] == [((x - centred_rem(x, ql)) // ql) % q for x in xs]


def centred_rem(x, q):
    r = x % q
    return r - q if r > q // 2 else r


def test_samplers(oracle):
    n = 4096
    s = oracle.sample_ternary_hw(n, 7, 0, 64)
    assert (s != 0).sum() == 64 and set(np.unique(s)) <= {-1, 0, 1}
    u = oracle.sample_ternary(n, 7, 0)
    assert set(np.unique(u)) == {-1, 0, 1} and abs((u != 0).mean() - 0.5) < 0.05
    cdt = oracle.gaussian_cdt(3.2)
    assert int(cdt[-1]) == 1 << 63 and all(int(cdt[i]) < int(cdt[i + 1]) for i in range(len(cdt) - 1))
    e = oracle.sample_gaussian(1 << 16, 7, 1, cdt)
    assert abs(e.mean()) < 0.1 and abs(e.std() - 3.2) < 0.1 and np.abs(e).max() <= 20
    # python restatement of the generator
    def mix(z):
        z &= (1 << 64) - 1
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & ((1 << 64) - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & ((1 << 64) - 1)
        return z ^ (z >> 31)
    G = 0x9E3779B97F4A7C15
    for seed, st, idx in [(0, 0, 0), (12345, 3, 99), (2**63, 1024, 2**40)]:
        assert oracle.rng64(seed, st, idx) == mix(mix(mix(seed) + G * (st + 1)) + G * (idx + 1))
        assert oracle.rng_key(seed, st) == mix(mix(seed) + G * (st + 1))
        assert oracle.item_seed(seed, st) == mix(mix(seed ^ 0x6A09E667F3BCC909) + G * (st + 1))
    q = 0xFFFFFFFFFFC0001
    a = oracle.sample_uniform(8, q, 5, 16)
    for j in range(8):
        assert int(a[j]) == ((oracle.rng64(5, 16, 2 * j) << 64) | oracle.rng64(5, 16, 2 * j + 1)) % q


@pytest.fixture(scope="module")
def small_bfv(oracle, chain):
    n, L, R, K, dnum, t = 32, 4, 5, 2, 2, 65537
    ctx = oracle.Bfv(n, L, R, K, dnum, t, chain[:L + R], sigma=3.2, hw=8)
    s, sk = ctx.secret_keygen(101)
    pk = ctx.public_keygen(102, sk)
    rlk = ctx.relin_keygen(103, sk)
    return ctx, s, sk, pk, rlk


def test_bfv_keys_are_rlwe_samples(oracle, chain, small_bfv):
    ctx, s, sk, pk, rlk = small_bfv
    n, L, K = ctx.n, ctx.L, ctx.K
    qs = chain[:L]
    Q = prod(qs)
    sl = [int(v) for v in s]
    # pk0 + pk1*s = e (small)
    pk0 = poly_from_rns([oracle.ntt_inverse(pk[0, i], qs[i]) for i in range(L)], qs)
    pk1 = poly_from_rns([oracle.ntt_inverse(pk[1, i], qs[i]) for i in range(L)], qs)
    e = [centred(a + b, Q) for a, b in zip(pk0, negacyclic_mul(pk1, sl, n))]
    assert max(abs(v) for v in e) <= 20
    # rlk_d: b + a*s - P*B_d*s^2 = e_d (small) mod QP
    W = chain[:L + K]
    QP = prod(W); P = prod(chain[L:L + K])
    s2 = negacyclic_mul(sl, sl, n)
    for d in range(ctx.dnum):
        grp = qs[d * ctx.alpha:(d + 1) * ctx.alpha]
        Qd = prod(grp); Qh = Q // Qd
        B = Qh * pow(Qh, -1, Qd)
        b = poly_from_rns([oracle.ntt_inverse(rlk[d, 0, i], W[i]) for i in range(L + K)], W)
        a = poly_from_rns([oracle.ntt_inverse(rlk[d, 1, i], W[i]) for i in range(L + K)], W)
        as_ = negacyclic_mul(a, sl, n)
        e = [centred(bb + aa - P * B * ss, QP) for bb, aa, ss in zip(b, as_, s2)]
        assert max(abs(v) for v in e) <= 20


def test_bfv_encrypt_decrypt_add(oracle, chain, small_bfv):
    ctx, s, sk, pk, rlk = small_bfv
    # reference examples: examples/basic_encryption.cu:46,90-106 ; tests/test_fhe.cu:201-202,264
    m = ctx.encode([42, 100, 255, 1337])
    ct = ctx.encrypt(201, m, pk)
    assert np.array_equal(ctx.decrypt(ct, sk), m)
    m1 = ctx.encode([5, 10, 15, 20]); m2 = ctx.encode([3, 6, 9, 12])
    c1 = ctx.encrypt(202, m1, pk); c2 = ctx.encrypt(203, m2, pk)
    assert [int(v) for v in ctx.decrypt(ctx.add(c1, c2), sk)[:4]] == [8, 16, 24, 32]
    # ciphertext really is (pk0 u + e1 + Delta m, pk1 u + e2): c0 + c1 s = Delta m + small
    qs = chain[:ctx.L]; Q = prod(qs)
    sl = [int(v) for v in s]
    c0 = poly_from_rns(ct[0], qs); c1p = poly_from_rns(ct[1], qs)
    x = [centred(a + b, Q) for a, b in zip(c0, negacyclic_mul(c1p, sl, ctx.n))]
    delta = Q // ctx.t
    noise = [centred(xx - delta * int(mm), Q) for xx, mm in zip(x, m)]
    assert max(abs(v) for v in noise) < 2**12
    assert [int(v) for v in ctx.consts()["delta"]] == [delta % q for q in qs]


def test_bfv_multiply_relin_exact(oracle, chain, small_bfv):
    ctx, s, sk, pk, rlk = small_bfv
    n, L, R, K, t = ctx.n, ctx.L, ctx.R, ctx.K, ctx.t
    qs = chain[:L]; Q = prod(qs)
    rng = np.random.default_rng(50)
    m1 = rng.integers(0, t, n, dtype=np.uint64); m2 = rng.integers(0, t, n, dtype=np.uint64)
    ca = ctx.encrypt(301, m1, pk); cb = ctx.encrypt(302, m2, pk)
    out, sc = ctx.multiply_relin(ca, cb, rlk, want_scaled=True)
    # decrypts to the negacyclic product mod t  (BASELINE.json config 2 check)
    expect = [v % t for v in negacyclic_mul([int(v) for v in m1], [int(v) for v in m2], n)]
    assert [int(v) for v in ctx.decrypt(out, sk)] == expect
    # step-exact: scaled tensor == round(t/Q * tensor over Z) reduced mod q_i
    a0, a1 = poly_from_rns(ca[0], qs), poly_from_rns(ca[1], qs)
    b0, b1 = poly_from_rns(cb[0], qs), poly_from_rns(cb[1], qs)
    d = [negacyclic_mul(a0, b0, n),
         [x + y for x, y in zip(negacyclic_mul(a0, b1, n), negacyclic_mul(a1, b0, n))],
         negacyclic_mul(a1, b1, n)]
    ds = [[round_half_up(t * v, Q) for v in poly] for poly in d]
    for p in range(3):
        for i, q in enumerate(qs):
            assert [int(v) for v in sc[p, i]] == [v % q for v in ds[p]]
    # step-exact: hybrid key switching with exact ModUp / ModDown
    W = chain[:L + K]; QP = prod(W); P = prod(chain[L:L + K])
    acc0 = [0] * n; acc1 = [0] * n
    for dg in range(ctx.dnum):
        grp = qs[dg * ctx.alpha:(dg + 1) * ctx.alpha]; Qd = prod(grp)
        D = [centred(v, Qd) for v in ds[2]]
        kb = poly_from_rns([oracle.ntt_inverse(rlk[dg, 0, i], W[i]) for i in range(L + K)], W, centre=False)
        ka = poly_from_rns([oracle.ntt_inverse(rlk[dg, 1, i], W[i]) for i in range(L + K)], W, centre=False)
        acc0 = [(x + y) % QP for x, y in zip(acc0, negacyclic_mul(D, kb, n))]
        acc1 = [(x + y) % QP for x, y in zip(acc1, negacyclic_mul(D, ka, n))]
    for p, acc in enumerate((acc0, acc1)):
        down = [(v - centred(v, P)) // P for v in acc]            # round(acc / P)
        for i, q in enumerate(qs):
            assert [int(v) for v in out[p, i]] == [(a + b) % q for a, b in zip(ds[p], down)]


def test_bfv_reference_style_products(oracle, chain):
    # BASELINE.json config 2 shape at reduced N: L=2 (log_q 120), t=65537
    n, L, R, K, dnum, t = 64, 2, 3, 1, 2, 65537
    ctx = oracle.Bfv(n, L, R, K, dnum, t, chain[:L + R], hw=16)
    s, sk = ctx.secret_keygen(1); pk = ctx.public_keygen(2, sk); rlk = ctx.relin_keygen(3, sk)
    m1 = ctx.encode([5, 10, 15, 20]); m2 = ctx.encode([3, 6, 9, 12])
    c = ctx.multiply_relin(ctx.encrypt(4, m1, pk), ctx.encrypt(5, m2, pk), rlk)
    # coefficient encoding => negacyclic polynomial product (15,60,150,300,375,360,240; SURVEY section 4 prints 355 for x^4, a typo: 10*12+15*9+20*6 = 375)
    assert [int(v) for v in ctx.decrypt(c, sk)[:8]] == [15, 60, 150, 300, 375, 360, 240, 0]
    # depth 2
    c2 = ctx.multiply_relin(c, ctx.encrypt(6, ctx.encode([2]), pk), rlk)
    assert [int(v) for v in ctx.decrypt(c2, sk)[:7]] == [30, 120, 300, 600, 750, 720, 480]


def test_bfv_plain_ops_and_batch_encoding(oracle, chain):
    """'next' rows: sub / add_plain / sub_plain / multiply_plain and SIMD slot encoding.  With slot encoding
    the reference's printed expectations become true: {5,10,15,20} x {3,6,9,12} -> 15 60 135 240 (tests/test_fhe.cu:270),
    {3,4,5,6} x {2,5,10,3} -> 6 20 50 18 and x pt{2} -> 20 40 60 80 (examples/homomorphic_operations.cu:148,242)."""
    n, L, R, K, dnum, t = 64, 2, 3, 1, 2, 65537
    ctx = oracle.Bfv(n, L, R, K, dnum, t, chain[:L + R], hw=16)
    _, sk = ctx.secret_keygen(1); pk = ctx.public_keygen(2, sk); rlk = ctx.relin_keygen(3, sk)
    rng = np.random.default_rng(70)
    m1 = rng.integers(0, t, n, dtype=np.uint64); m2 = rng.integers(0, t, n, dtype=np.uint64)
    c1 = ctx.encrypt(4, m1, pk); c2 = ctx.encrypt(5, m2, pk)
    tt = np.uint64(t)
    assert np.array_equal(ctx.decrypt(ctx.sub(c1, c2), sk), (m1 + tt - m2) % tt)
    assert np.array_equal(ctx.decrypt(ctx.add_plain(c1, m2), sk), (m1 + m2) % tt)
    assert np.array_equal(ctx.decrypt(ctx.add_plain(c1, m2, subtract=True), sk), (m1 + tt - m2) % tt)
    assert np.array_equal(ctx.decrypt(ctx.multiply_plain(c1, m2), sk), oracle.schoolbook_negacyclic(m1, m2, t))
    # slot encoding
    e = lambda v: ctx.encrypt(6 + sum(v), ctx.batch_encode(v), pk)
    d = lambda c: [int(x) for x in ctx.batch_decode(ctx.decrypt(c, sk))[:4]]
    assert d(e([5, 10, 15, 20])) == [5, 10, 15, 20]
    assert d(ctx.multiply_relin(e([5, 10, 15, 20]), e([3, 6, 9, 12]), rlk)) == [15, 60, 135, 240]
    assert d(ctx.multiply_relin(e([3, 4, 5, 6]), e([2, 5, 10, 3]), rlk)) == [6, 20, 50, 18]
    assert d(ctx.add(e([10, 20, 30, 40]), e([5, 15, 25, 35]))) == [15, 35, 55, 75]
    assert d(ctx.multiply_plain(e([10, 20, 30, 40]), ctx.batch_encode([2] * n))) == [20, 40, 60, 80]


def _galois_ref(a, g, q):
    """a(x) -> a(x^g) mod (x^n + 1, q), written out from the definition."""
    n = len(a); out = [0] * n
    for i, v in enumerate(a):
        e = (i * g) % (2 * n)
        if e < n: out[e] = (out[e] + int(v)) % q
        else: out[e - n] = (out[e - n] - int(v)) % q
    return out


def test_galois_keys_rotation_and_mod_switch(oracle, chain, small_bfv):
    """'next' rows: Galois keys + automorphisms (rotate_rows / rotate_columns, include/fhe.cuh:86,113-116)
    and mod_switch_to_next (include/fhe.cuh:109).  Pinned by definition: the automorphism against x -> x^g written out, the key
    as an RLWE sample of P*B_d*s(x^g), decrypt(apply_galois(enc m)) = m(x^g), the slot permutation k -> k' with
    (2 bitrev(k') + 1) = (2 bitrev(k) + 1) g mod 2N under batch encoding, and decryption under the (L-1)-limb context."""
    ctx, s, sk, pk, rlk = small_bfv
    n, L, K, t = ctx.n, ctx.L, ctx.K, ctx.t
    qs = chain[:L]; Q = prod(qs)
    sl = [int(v) for v in s]
    rng = np.random.default_rng(71)
    a = rng.integers(0, qs[0], n, dtype=np.uint64)
    for g in (3, 9, 2 * n - 1, 5, 27):
        assert [int(v) for v in oracle.apply_galois_poly(a, g, qs[0])] == _galois_ref(a, g, qs[0])
    W = chain[:L + K]; QP = prod(W); P = prod(chain[L:L + K])
    m = rng.integers(0, t, n, dtype=np.uint64)
    ct = ctx.encrypt(301, m, pk)
    lg = n.bit_length() - 1
    brev = lambda k: int(format(k, f"0{lg}b")[::-1], 2)
    for g in (3, 2 * n - 1, 3 ** 5 % (2 * n)):
        gk = ctx.galois_keygen(400 + g, g, sk)
        # gk_d: b + a*s - P*B_d*s(x^g) = e_d (small) mod QP
        sg = [centred(v, Q * P) for v in _galois_ref([v % QP for v in sl], g, QP)]
        for d in range(ctx.dnum):
            grp = qs[d * ctx.alpha:(d + 1) * ctx.alpha]
            Qd = prod(grp); Qh = Q // Qd
            B = Qh * pow(Qh, -1, Qd)
            b = poly_from_rns([oracle.ntt_inverse(gk[d, 0, i], W[i]) for i in range(L + K)], W)
            aa = poly_from_rns([oracle.ntt_inverse(gk[d, 1, i], W[i]) for i in range(L + K)], W)
            e = [centred(bb + x - P * B * ss, QP) for bb, x, ss in zip(b, negacyclic_mul(aa, sl, n), sg)]
            assert max(abs(v) for v in e) <= 20
        rot = ctx.apply_galois(ct, g, gk)
        assert [int(v) for v in ctx.decrypt(rot, sk)] == _galois_ref(m, g, t)
        # the rotated ciphertext is a fresh-looking encryption under s: c0 + c1 s = Delta m(x^g) + small
        c0 = poly_from_rns(rot[0], qs); c1 = poly_from_rns(rot[1], qs)
        x = [centred(u + v, Q) for u, v in zip(c0, negacyclic_mul(c1, sl, n))]
        noise = [centred(xx - (Q // t) * mm, Q) for xx, mm in zip(x, _galois_ref(m, g, t))]
        assert max(abs(v) for v in noise) < 2**40
        # slots (t = 65537 = 1 mod 2N): decode(m(x^g))[k] = decode(m)[k'] with exponent(k') = exponent(k) * g
        slots = ctx.batch_decode(m); rslots = ctx.batch_decode(ctx.decrypt(rot, sk))
        inv = {2 * brev(k) + 1: k for k in range(n)}
        for k in range(n):
            assert rslots[k] == slots[inv[((2 * brev(k) + 1) * g) % (2 * n)]]
    # modulus chain: drop q_{L-1} with rounding; same plaintext under the context on the first L-1 limbs (same secret seed)
    low = oracle.Bfv(n, L - 1, ctx.R, K, 1, t, list(chain[:L - 1]) + list(chain[L:L + ctx.R]), sigma=3.2, hw=8)
    s2, sk2 = low.secret_keygen(101)
    assert np.array_equal(s2, s)
    ct2 = ctx.mod_switch_to_next(ct)
    assert ct2.shape == (2, L - 1, n)
    assert np.array_equal(low.decrypt(ct2, sk2), m)
    for p in range(2):
        assert np.array_equal(ct2[p], oracle.modswitch_drop_last(ct[p], qs))


def test_multiply_then_relinearize_is_the_fused_multiply(oracle):
    """orc_bfv_multiply (3 components) + orc_bfv_relinearize == orc_bfv_multiply_relin; the 3-component ciphertext decrypts with
    (1, s, s^2) to the product, and a sum of two products relinearised once decrypts to the sum of the products."""
    from fhe_b200.params import bfv_preset
    p = bfv_preset("small")
    o = oracle.Bfv(p["n"], p["L"], p["R"], p["K"], p["dnum"], p["t"], p["primes"], sigma=p["sigma"], hw=p["hamming_weight"])
    n, t, L = p["n"], p["t"], p["L"]
    _, sk = o.secret_keygen(5); pk = o.public_keygen(6, sk); rlk = o.relin_keygen(7, sk)
    rng = np.random.default_rng(77)
    m = rng.integers(0, t, (4, n), dtype=np.uint64)
    ct = [o.encrypt(100 + i, m[i], pk) for i in range(4)]
    p01, p23 = o.multiply(ct[0], ct[1]), o.multiply(ct[2], ct[3])
    fused, scaled = o.multiply_relin(ct[0], ct[1], rlk, want_scaled=True)
    assert np.array_equal(p01, scaled)
    assert np.array_equal(o.relinearize(p01, rlk), fused)
    mods = np.array(p["primes"][:L], dtype=np.uint64)[None, :, None]
    lazy = o.relinearize((p01 + p23) % mods, rlk)
    want = (oracle.schoolbook_negacyclic(m[0], m[1], t) + oracle.schoolbook_negacyclic(m[2], m[3], t)) % np.uint64(t)
    assert np.array_equal(o.decrypt(lazy, sk), want)


def test_mod_switch_to_level_is_repeated_drop_and_still_decrypts(oracle):
    """orc_bfv_mod_switch_to_level(drop) == drop applications of the one-limb switch (each on the shorter chain), and the result
    decrypts under the context built on the remaining limbs."""
    from fhe_b200.params import bfv_preset
    p = bfv_preset("small")
    n, t, L = p["n"], p["t"], p["L"]
    o = oracle.Bfv(n, L, p["R"], p["K"], p["dnum"], t, p["primes"], sigma=p["sigma"], hw=p["hamming_weight"])
    _, sk = o.secret_keygen(5); pk = o.public_keygen(6, sk)
    m = np.random.default_rng(9).integers(0, t, n, dtype=np.uint64)
    ct = o.encrypt(7, m, pk)
    assert np.array_equal(o.mod_switch_to_level(ct, 0), ct)
    assert np.array_equal(o.mod_switch_to_level(ct, 1), o.mod_switch_to_next(ct))
    for drop in (2, 3):
        low = o.mod_switch_to_level(ct, drop)
        assert low.shape == (2, L - drop, n)
        step = ct
        for k in range(drop):                      # one limb at a time through contexts of decreasing length
            primes_k = list(p["primes"][:L - k]) + list(p["primes"][L:])
            ok = oracle.Bfv(n, L - k, p["R"], p["K"], 1, t, primes_k, sigma=p["sigma"], hw=p["hamming_weight"])
            step = ok.mod_switch_to_next(step)
        assert np.array_equal(low, step)
        primes = list(p["primes"][:L - drop]) + list(p["primes"][L:])
        lo = oracle.Bfv(n, L - drop, p["R"], p["K"], 1, t, primes, sigma=p["sigma"], hw=p["hamming_weight"])
        _, lsk = lo.secret_keygen(5)
        assert np.array_equal(lo.decrypt(low, lsk), m)


def test_generator_keys_never_alias_across_calls_items_and_streams(oracle):
    """with key = mix64(seed + G (stream + 1)) a caller stepping its seed by G (the compat layer did) made
    (seed_k, stream s) collide with (seed_k+1, stream s-1), and encrypt's seed + b made consecutive seeds overlap.  The keys are now
    mix64(mix64(seed) + G (stream + 1)) with batch items in a domain of their own: every (call, item, stream) of a
    keygen / relinkeygen / galoiskeygen / encrypt x k sequence must get its own key -- for G-spaced seeds and for consecutive seeds."""
    G = 0x9E3779B97F4A7C15
    L, W, dnum = 24, 32, 3
    for step in (G, 1):
        keys = {}

        def use(tag, seed, stream):
            k = oracle.rng_key(seed, stream)
            assert k not in keys, (tag, keys[k])
            keys[k] = tag

        seed = 12345
        calls = []
        for call in range(40):
            seed = (seed + step) & (2**64 - 1)
            calls.append(seed)
        it = iter(calls)
        s_sk, s_pk, s_rlk, s_gk = next(it), next(it), next(it), next(it)
        for st in (0, 1):
            use(("sk", st), s_sk, st)
        use(("pk_e",), s_pk, 2)
        for i in range(L):
            use(("pk_a", i), s_pk, 16 + i)
        for name, sd in (("rlk", s_rlk), ("gk", s_gk)):
            for d in range(dnum):
                use((name, "e", d), sd, 1024 * (d + 1) + 512)
                for i in range(W):
                    use((name, "a", d, i), sd, 1024 * (d + 1) + i)
        for call, sd in enumerate(it):                     # encrypt calls of batch 8
            for b in range(8):
                for st in (0, 1, 2):
                    use(("enc", call, b, st), oracle.item_seed(sd, b), st)
    # the old derivation would have failed this: (seed, stream s) == (seed + G, stream s - 1)
    assert oracle.rng_key(7, 1) != oracle.rng_key((7 + G) & (2**64 - 1), 0)


def test_chacha20_block_known_answer_and_keyed_samplers(oracle):
    """the keyed generator (mirror of fhe_b200_bfv_set_rng_key): RFC 8439 section 2.3.2 block test vector, and the samplers switch to
    it and back"""
    key = [int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)]
    out = oracle.chacha20_block(key, 1, [0x09000000, 0x4A000000, 0x00000000])
    assert out == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
                   0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]
    plain = oracle.sample_ternary(256, 7, 0)
    try:
        oracle.set_rng_key(bytes(range(32)))
        # nonce (seed low, seed high, stream) = (0x09000000, 0x4a000000, 0), counter 1 = words 8..15
        seed = 0x4A00000009000000
        assert [oracle.rng64(seed, 0, 8 + w) for w in range(8)] == [(out[2 * w + 1] << 32) | out[2 * w] for w in range(8)]
        keyed = oracle.sample_ternary(256, 7, 0)
        assert not np.array_equal(keyed, plain) and set(np.unique(keyed)) <= {-1, 0, 1}
    finally:
        oracle.set_rng_key(None)
    assert np.array_equal(oracle.sample_ternary(256, 7, 0), plain)
