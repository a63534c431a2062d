"""GPU parity tests for the RNS layer and the BFV path (run with -m gpu): CUDA through the C ABI vs the CPU oracle,
bit-exact on every ciphertext word."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def fhe():
    import fhe_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    fhe_b200.load_library()
    return fhe_b200


def _rand_limbs(rng, mods, n, batch):
    return np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods]) for _ in range(batch)])


def _pick_lincomb_kernel(monkeypatch, tpc):
    """IMAD kernel with one or two lanes per coefficient ("1", "2"), the mma.sync tensor-core kernel ("mma") or the tcgen05 / TMEM
    kernel ("tc": eight folded columns per target where every target is a chain prime; "tcb": its Toeplitz form with the 128-bit Barrett
    epilogue, the one other moduli get; shapes whose operands do not fit shared memory fall back to mma.sync) -- read when the object is created"""
    monkeypatch.setenv("FHE_B200_LINCOMB_MMA", "1" if tpc in ("mma", "tc", "tcb") else "0")
    monkeypatch.setenv("FHE_B200_LINCOMB_TC", "1" if tpc in ("tc", "tcb") else "0")
    monkeypatch.setenv("FHE_B200_LINCOMB_TC_FOLD", "0" if tpc == "tcb" else "1")
    if tpc in ("1", "2"):
        monkeypatch.setenv("FHE_B200_LINCOMB_TPC", tpc)


@pytest.mark.parametrize("tpc", ["1", "2", "mma", "tc", "tcb"])
@pytest.mark.parametrize("S,T", [(1, 3), (2, 3), (4, 5), (8, 24), (24, 25), (25, 24), (9, 4), (30, 7), (33, 2), (49, 25), (62, 3)])
def test_base_conversion_vs_oracle(fhe, oracle, chain, S, T, tpc, monkeypatch):
    from fhe_b200.engine import to_device, to_host
    _pick_lincomb_kernel(monkeypatch, tpc)
    src, dst = chain[:S], chain[S:S + T]
    n, batch = 256, 3
    rng = np.random.default_rng(100 + S)
    x = _rand_limbs(rng, src, n, batch)
    x[0, :, 0] = 0; x[0, :, 1] = [q - 1 for q in src]; x[0, :, 2] = 1
    lc = fhe.LinComb.conv(src, dst)
    olc = oracle.LinComb.conv(src, dst)
    got = to_host(lc.apply(to_device(x)))
    for b in range(batch):
        assert np.array_equal(got[b], olc.apply(x[b])), b
    cg, co = lc.constants(), olc.constants()
    for k in co:
        assert np.array_equal(cg[k], co[k]), k


@pytest.mark.parametrize("tpc", ["1", "2", "mma", "tc", "tcb"])
@pytest.mark.parametrize("L,R,t", [(1, 2, 65537), (2, 3, 65537), (4, 5, 786433), (24, 25, 65537), (3, 4, 1 << 20)])
def test_scale_and_round_vs_oracle(fhe, oracle, chain, L, R, t, tpc, monkeypatch):
    from fhe_b200.engine import to_device, to_host
    _pick_lincomb_kernel(monkeypatch, tpc)
    qs, ps = chain[:L], chain[L:L + R]
    n, batch = 512, 2
    rng = np.random.default_rng(200 + L)
    dq = _rand_limbs(rng, qs, n, batch); dp = _rand_limbs(rng, ps, n, batch)
    lc = fhe.LinComb.scale(qs, ps, t, ps, True)
    olc = oracle.LinComb.scale(qs, ps, t, ps, True)
    got = to_host(lc.apply(to_device(dq), to_device(dp)))
    for b in range(batch):
        assert np.array_equal(got[b], olc.apply(dq[b], extra=dp[b]))
    # decryption scaling: targets = {t}
    lcd = fhe.LinComb.scale(qs, [], t, [t], False)
    old = oracle.LinComb.scale(qs, [], t, [t], False)
    got = to_host(lcd.apply(to_device(dq)))
    for b in range(batch):
        assert np.array_equal(got[b], old.apply(dq[b]))


def test_modswitch_drop_last_vs_oracle(fhe, oracle, chain):
    from fhe_b200.engine import to_device, to_host
    n, mods = 1024, chain[:4]
    plan = fhe.Plan(n, mods)
    rng = np.random.default_rng(300)
    x = _rand_limbs(rng, mods, n, 2)
    x[0, 3, :4] = [0, mods[3] // 2, mods[3] // 2 + 1, mods[3] - 1]
    got = to_host(plan.modswitch_drop_last(to_device(x)))
    for b in range(2):
        assert np.array_equal(got[b], oracle.modswitch_drop_last(x[b], mods))


def test_gaussian_cdt_matches_oracle(fhe, oracle):
    for sigma in (3.2, 3.19, 1.0, 8.0):
        assert np.array_equal(fhe.gaussian_cdt(sigma), oracle.gaussian_cdt(sigma))


def _exact_product_mod_t(oracle, a, b, t, q):
    """a*b mod (x^N+1, t) for inputs < t < 2^17: the integer product has |coeff| < N t^2 < 2^51 < q/2, so the
    negacyclic product modulo the 60-bit NTT prime q, lifted to the centred range, is the integer product."""
    r = oracle.negacyclic_mul_ntt(a, b, q)
    signed = np.where(r > np.uint64(q // 2), r.astype(np.int64) - np.int64(q), r.astype(np.int64))
    return np.mod(signed, np.int64(t)).astype(np.uint64)


def _setup(fhe, oracle, preset, hw=None):
    from fhe_b200.params import bfv_preset
    p = bfv_preset(preset)
    if hw is not None:
        p["hamming_weight"] = hw
    g = fhe.BfvContext(p["n"], p["L"], p["R"], p["K"], p["dnum"], p["t"], p["primes"], p["sigma"], p["hamming_weight"])
    o = oracle.Bfv(p["n"], p["L"], p["R"], p["K"], p["dnum"], p["t"], p["primes"], sigma=p["sigma"], hw=p["hamming_weight"])
    return p, g, o


@pytest.mark.parametrize("preset,hw", [("small", 64), ("small", 0), ("c2", 64)])
def test_bfv_keys_encrypt_decrypt_add_vs_oracle(fhe, oracle, preset, hw):
    from fhe_b200.engine import to_device, to_host
    p, g, o = _setup(fhe, oracle, preset, hw)
    n, t = p["n"], p["t"]
    sk, pk = g.keygen(11, 12)
    s_small, osk = o.secret_keygen(11)
    opk = o.public_keygen(12, osk)
    assert np.array_equal(to_host(sk), osk)
    assert np.array_equal(to_host(pk), opk)
    rlk = g.relinkey_gen(13, sk)
    assert np.array_equal(to_host(rlk), o.relin_keygen(13, osk))
    rng = np.random.default_rng(400)
    m = rng.integers(0, t, (3, n), dtype=np.uint64)
    m[0, :4] = [5, 10, 15, 20]; m[0, 4:] = 0            # tests/test_fhe.cu:201
    m[1, :4] = [3, 6, 9, 12]; m[1, 4:] = 0              # tests/test_fhe.cu:202
    ct = g.encrypt(500, to_device(m), pk)               # ciphertext b is batch item b of the call seeded 500
    h = to_host(ct)
    for b in range(3):
        assert np.array_equal(h[b], o.encrypt(500, m[b], opk, item=b)), b
    assert np.array_equal(to_host(g.decrypt(ct, sk)), m)
    s = g.add(ct[0:1].contiguous(), ct[1:2].contiguous())
    assert np.array_equal(to_host(s)[0], o.add(h[0], h[1]))
    assert [int(v) for v in to_host(g.decrypt(s, sk))[0, :4]] == [8, 16, 24, 32]      # tests/test_fhe.cu:264
    # reference example: encrypt -> decrypt identity on {42,100,255,1337} (examples/basic_encryption.cu:46,90-106)
    pt = g.encode([42, 100, 255, 1337])
    assert [int(v) for v in g.decode(g.decrypt(g.encrypt(7, pt, pk), sk))[0, :4]] == [42, 100, 255, 1337]


@pytest.mark.parametrize("preset", ["small", "c2", "mid", "mid-fused"])
def test_bfv_multiply_relin_vs_oracle(fhe, oracle, preset, monkeypatch):
    """BASELINE.json config 2 (preset c2): every ciphertext word equals the oracle's; decrypts to the negacyclic product.
    Preset mid (N = 8192) runs the two-pass transform; mid-fused the same with the fused tile kernels (tensor product and key-switch
    inner product inside the tile passes, csrc/ntt_fused.cu, FHE_B200_FUSED_TILE=1)."""
    from fhe_b200.engine import to_device, to_host
    if preset.endswith("-fused"):
        monkeypatch.setenv("FHE_B200_FUSED_TILE", "1")
        preset = preset[:-6]
    p, g, o = _setup(fhe, oracle, preset)
    n, t = p["n"], p["t"]
    sk, pk = g.keygen(21, 22); rlk = g.relinkey_gen(23, sk)
    _, osk = o.secret_keygen(21); orlk = o.relin_keygen(23, osk)
    rng = np.random.default_rng(600)
    m1 = rng.integers(0, t, (2, n), dtype=np.uint64); m2 = rng.integers(0, t, (2, n), dtype=np.uint64)
    m1[0] = 0; m1[0, :4] = [5, 10, 15, 20]; m2[0] = 0; m2[0, :4] = [3, 6, 9, 12]
    ca = g.encrypt(700, to_device(m1), pk); cb = g.encrypt(800, to_device(m2), pk)
    out, sc = g.multiply(ca, cb, rlk, want_scaled=True)
    ha, hb, ho, hs = to_host(ca), to_host(cb), to_host(out), to_host(sc)
    for b in range(2):
        eo, es = o.multiply_relin(ha[b], hb[b], orlk, want_scaled=True)
        assert np.array_equal(hs[b], es), ("scaled tensor", b)
        assert np.array_equal(ho[b], eo), ("relinearised", b)
    dec = to_host(g.decrypt(out, sk))
    assert [int(v) for v in dec[0, :8]] == [15, 60, 150, 300, 375, 360, 240, 0]
    assert np.array_equal(dec[1], oracle.schoolbook_negacyclic(m1[1], m2[1], t))
    # depth 2
    out2 = g.multiply(out, ca, rlk)
    exp = oracle.schoolbook_negacyclic(oracle.schoolbook_negacyclic(m1[1], m2[1], t), m1[1], t)
    assert np.array_equal(to_host(g.decrypt(out2, sk))[1], exp)
    # host-buffer entry point gives the same words
    h_out = np.zeros_like(ha)
    g.multiply_host(ha, hb, rlk, h_out)
    assert np.array_equal(h_out, ho)


@pytest.mark.slow
def test_bfv_config4_properties(fhe, oracle):
    """BASELINE.json config 4 at full size (N=2^16, L=24, R=25, dnum=3, K=8).  The oracle needs minutes here, so the
    check is by properties: decrypt(HMult(enc m1, enc m2)) == m1*m2 mod (x^N+1, t) (host NTT product), add is
    homomorphic, and two limbs of the key material are spot-checked against the oracle's samplers."""
    from fhe_b200.engine import to_device, to_host
    from fhe_b200.params import bfv_preset
    p = bfv_preset("c4")
    n, t = p["n"], p["t"]
    g = fhe.BfvContext(n, p["L"], p["R"], p["K"], p["dnum"], t, p["primes"], p["sigma"], p["hamming_weight"])
    sk, pk = g.keygen(31, 32); rlk = g.relinkey_gen(33, sk)
    q0 = p["primes"][0]
    assert np.array_equal(to_host(pk)[1, 0], oracle.sample_uniform(n, q0, 32, 16))
    rng = np.random.default_rng(900)
    m1 = rng.integers(0, t, n, dtype=np.uint64); m2 = rng.integers(0, t, n, dtype=np.uint64)
    ca = g.encrypt(41, to_device(m1.reshape(1, n)), pk); cb = g.encrypt(42, to_device(m2.reshape(1, n)), pk)
    assert np.array_equal(to_host(g.decrypt(ca, sk))[0], m1)
    out = g.multiply(ca, cb, rlk)
    m12 = _exact_product_mod_t(oracle, m1, m2, t, q0)
    assert np.array_equal(to_host(g.decrypt(out, sk))[0], m12)
    out2 = g.multiply(out, cb, rlk)
    assert np.array_equal(to_host(g.decrypt(out2, sk))[0], _exact_product_mod_t(oracle, m12, m2, t, q0))
    s = g.add(ca, cb)
    assert np.array_equal(to_host(g.decrypt(s, sk))[0], (m1 + m2) % np.uint64(t))


def test_bfv_config4_every_word_vs_oracle_and_committed_digests(fhe, oracle):
    """BASELINE.json config 4 at FULL size, word for word: keys, a ciphertext, the scaled tensor and the relinearised product equal
    the oracle's (run here, ~10 s of host time) and the digests committed in tests/golden/golden_slow.json; the squaring path and
    the separate (unfused) tensor / inner-product kernels give the same words."""
    import hashlib, json, os
    from fhe_b200.engine import to_device, to_host
    p, g, o = _setup(fhe, oracle, "c4")
    n, t = p["n"], p["t"]
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_slow.json")))
    dg = lambda a: hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()
    sk, pk = g.keygen(31, 32); rlk = g.relinkey_gen(33, sk)
    _, osk = o.secret_keygen(31); opk = o.public_keygen(32, osk); orlk = o.relin_keygen(33, osk)
    assert np.array_equal(to_host(sk), osk) and np.array_equal(to_host(pk), opk) and np.array_equal(to_host(rlk), orlk)
    assert dg(osk) == gold["bfv_c4/sk"]["output"] and dg(opk) == gold["bfv_c4/pk"]["output"] and dg(orlk) == gold["bfv_c4/rlk"]["output"]
    m1 = np.random.default_rng(34).integers(0, t, n, dtype=np.uint64); m2 = np.random.default_rng(35).integers(0, t, n, dtype=np.uint64)
    c1 = g.encrypt(36, to_device(m1[None]), pk); c2 = g.encrypt(37, to_device(m2[None]), pk)
    h1, h2 = to_host(c1)[0], to_host(c2)[0]
    assert np.array_equal(h1, o.encrypt(36, m1, opk)) and dg(h1) == gold["bfv_c4/encrypt"]["output"]
    out, sc = g.multiply(c1, c2, rlk, want_scaled=True)
    eo, es = o.multiply_relin(h1, h2, orlk, want_scaled=True)
    assert np.array_equal(to_host(sc)[0], es) and dg(es) == gold["bfv_c4/scaled_tensor"]["output"]
    assert np.array_equal(to_host(out)[0], eo) and dg(eo) == gold["bfv_c4/multiply_relin"]["output"]
    assert dg(to_host(g.multiply(c1, c1, rlk))[0]) == gold["bfv_c4/square"]["output"]
    # batch of three (halves of 2 and 1 on two streams) gives the same words per ciphertext
    c3a = torch.cat([c1, c2, c1]); c3b = torch.cat([c2, c1, c1])
    o3 = to_host(g.multiply(c3a, c3b, rlk))
    assert np.array_equal(o3[0], eo) and np.array_equal(o3[1], eo) and dg(o3[2]) == gold["bfv_c4/square"]["output"]
    lib = fhe.load_library()
    l0 = lib.fhe_b200_launch_count()
    out_sep = g.multiply(c1, c2, rlk)
    sep = lib.fhe_b200_launch_count() - l0
    os.environ["FHE_B200_FUSED_TILE"] = "1"            # tensor product and inner product inside the tile passes (read at every call)
    try:
        l0 = lib.fhe_b200_launch_count()
        out_fused = g.multiply(c1, c2, rlk)
        fused = lib.fhe_b200_launch_count() - l0
        assert dg(to_host(g.multiply(c1, c1, rlk))[0]) == gold["bfv_c4/square"]["output"]
        o3f = to_host(g.multiply(c3a, c3b, rlk))
    finally:
        del os.environ["FHE_B200_FUSED_TILE"]
    assert torch.equal(out_sep, out) and torch.equal(out_fused, out) and np.array_equal(o3f, o3)
    assert fused < sep, (fused, sep)


def test_limb_sharded_building_blocks_single_rank(fhe, oracle):
    """parallel.LimbShardedBfv with the GPU backend at world=1 must reproduce fhe_b200_bfv_multiply_relin bit for bit
    (exercises fhe_b200_bfv_tensor / _ks_inner and the subset-target scale constants on the device)."""
    from fhe_b200.engine import to_device, to_host
    import fhe_b200.parallel as par
    p, g, o = _setup(fhe, oracle, "small")
    n, t = p["n"], p["t"]
    sk, pk = g.keygen(51, 52); rlk = g.relinkey_gen(53, sk)
    rng = np.random.default_rng(54)
    ca = g.encrypt(55, to_device(rng.integers(0, t, (1, n), dtype=np.uint64)), pk)
    cb = g.encrypt(56, to_device(rng.integers(0, t, (1, n), dtype=np.uint64)), pk)
    expect = to_host(g.multiply(ca, cb, rlk))[0]
    be = par.GpuBackend(n, p["primes"], 0)
    sb = par.LimbShardedBfv(n, p["L"], p["R"], p["K"], p["dnum"], t, p["primes"], be, rank=0, world=1)
    kq, kp = sb.shard_relin_key(rlk)
    out = sb.multiply_relin(ca[0].contiguous(), cb[0].contiguous(), kq, kp)
    assert np.array_equal(to_host(out), expect)
    # emulate a 2-rank split on one GPU: each "rank" computes its limbs from the same gathered inputs
    for world in (2, 3):
        for r in range(world):
            sh = par.LimbShard(p["L"], p["R"], p["K"], r, world)
            lc = fhe.LinComb.scale(p["primes"][:p["L"]], p["primes"][p["L"]:], t, p["primes"][p["L"] + sh.rb:p["L"] + sh.rb + sh.rc], True)
            oc = oracle.LinComb.scale(p["primes"][:p["L"]], p["primes"][p["L"]:], t, p["primes"][p["L"] + sh.rb:p["L"] + sh.rb + sh.rc], True)
            if sh.rc == 0:
                continue
            dq = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in p["primes"][:p["L"]]])
            dp = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in p["primes"][p["L"] + sh.rb:p["L"] + sh.rb + sh.rc]])
            assert np.array_equal(to_host(lc.apply(to_device(dq[None]), to_device(dp[None])))[0], oc.apply(dq, extra=dp))


def test_plain_ops_and_batch_encoding_vs_oracle(fhe, oracle):
    from fhe_b200.engine import to_device, to_host
    p, g, o = _setup(fhe, oracle, "small")
    n, t = p["n"], p["t"]
    sk, pk = g.keygen(61, 62); rlk = g.relinkey_gen(63, sk)
    _, osk = o.secret_keygen(61); opk = o.public_keygen(62, osk)
    rng = np.random.default_rng(64)
    m1 = rng.integers(0, t, (1, n), dtype=np.uint64); m2 = rng.integers(0, t, (1, n), dtype=np.uint64)
    c1 = g.encrypt(65, to_device(m1), pk); c2 = g.encrypt(66, to_device(m2), pk)
    h1, h2 = to_host(c1)[0], to_host(c2)[0]
    assert np.array_equal(to_host(g.sub(c1, c2))[0], o.sub(h1, h2))
    assert np.array_equal(to_host(g.add_plain(c1, to_device(m2)))[0], o.add_plain(h1, m2[0]))
    assert np.array_equal(to_host(g.add_plain(c1, to_device(m2), subtract=True))[0], o.add_plain(h1, m2[0], subtract=True))
    mp = g.multiply_plain(c1, to_device(m2))
    assert np.array_equal(to_host(mp)[0], o.multiply_plain(h1, m2[0]))
    c3 = torch.cat([c1, c2, c1]); m3 = np.concatenate([m2, m1, m1])                       # batch 3: halves of 2 and 1
    mp3 = to_host(g.multiply_plain(c3, to_device(m3)))
    for b, (hh, mm) in enumerate(((h1, m2[0]), (h2, m1[0]), (h1, m1[0]))):
        assert np.array_equal(mp3[b], o.multiply_plain(hh, mm)), b
    assert np.array_equal(to_host(g.decrypt(mp, sk))[0], oracle.schoolbook_negacyclic(m1[0], m2[0], t))
    # slot encoding: the reference's printed slot-wise expectations (tests/test_fhe.cu:270)
    pa, pb = g.batch_encode([5, 10, 15, 20]), g.batch_encode([3, 6, 9, 12])
    assert np.array_equal(to_host(pa), o.batch_encode([5, 10, 15, 20]))
    prod = g.multiply(g.encrypt(67, pa.view(1, n), pk), g.encrypt(68, pb.view(1, n), pk), rlk)
    assert [int(v) for v in g.batch_decode(g.decrypt(prod, sk))[0, :4]] == [15, 60, 135, 240]


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["small", "c2", "mid"])
def test_galois_and_mod_switch_vs_oracle(fhe, oracle, preset):
    """'next' rows: Galois keys, automorphisms (rotate_rows / rotate_columns) and mod_switch_to_next, bit for bit
    against the oracle (keys, rotated ciphertexts, switched ciphertexts) and by definition after decryption."""
    from fhe_b200.engine import to_device, to_host
    p, g, o = _setup(fhe, oracle, preset)
    n, t, L = p["n"], p["t"], p["L"]
    sk, pk = g.keygen(71, 72)
    _, osk = o.secret_keygen(71)
    rng = np.random.default_rng(73)
    m = rng.integers(0, t, (2, n), dtype=np.uint64)
    ct = g.encrypt(74, to_device(m), pk)
    h = to_host(ct)
    for elt in (3, 2 * n - 1, pow(3, 7, 2 * n)):
        gk = g.galoiskey_gen(80 + elt, elt, sk)
        ogk = o.galois_keygen(80 + elt, elt, osk)
        assert np.array_equal(to_host(gk), ogk), elt
        rot = g.apply_galois(ct, elt, gk)
        hr = to_host(rot)
        for b in range(2):
            assert np.array_equal(hr[b], o.apply_galois(h[b], elt, ogk)), (elt, b)
            assert np.array_equal(to_host(g.decrypt(rot, sk))[b], oracle.apply_galois_poly(m[b], elt, t)), (elt, b)
    # helpers: rotate_rows(steps) = x -> x^(3^steps), rotate_columns = x -> x^(2N-1)
    gk1 = g.galoiskey_gen(90, 3, sk)
    assert torch.equal(g.rotate_rows(ct, 1, gk1), g.apply_galois(ct, 3, gk1))
    # modulus chain
    low = to_host(g.mod_switch_to_next(ct))
    assert low.shape == (2, 2, L - 1, n)
    for b in range(2):
        assert np.array_equal(low[b], o.mod_switch_to_next(h[b])), b
    if L - 1 >= 1:
        primes = list(p["primes"][:L - 1]) + list(p["primes"][L:])
        lo = oracle.Bfv(n, L - 1, p["R"], p["K"], 1, t, primes, sigma=p["sigma"], hw=p["hamming_weight"])
        _, lsk = lo.secret_keygen(71)
        for b in range(2):
            assert np.array_equal(lo.decrypt(low[b], lsk), m[b]), b


@pytest.mark.gpu
def test_galois_config4_properties(fhe, oracle):
    """config 4 size: decrypt(apply_galois(enc m, g)) = m(x^g) for a row rotation and the column swap; rotations compose."""
    from fhe_b200.engine import to_device, to_host
    from fhe_b200.params import bfv_preset
    p = bfv_preset("c4"); n, t = p["n"], p["t"]
    g = fhe.BfvContext(n, p["L"], p["R"], p["K"], p["dnum"], t, p["primes"], p["sigma"], p["hamming_weight"])
    sk, pk = g.keygen(1, 2)
    rng = np.random.default_rng(9)
    m = rng.integers(0, t, (1, n), dtype=np.uint64)
    ct = g.encrypt(3, to_device(m), pk)
    g3 = g.galoiskey_gen(4, 3, sk); gc = g.galoiskey_gen(5, 2 * n - 1, sk)
    r = g.rotate_rows(ct, 1, g3)
    assert np.array_equal(to_host(g.decrypt(r, sk))[0], oracle.apply_galois_poly(m[0], 3, t))
    rr = g.rotate_rows(r, 1, g3)
    assert np.array_equal(to_host(g.decrypt(rr, sk))[0], oracle.apply_galois_poly(m[0], 9, t))
    c = g.rotate_columns(g.rotate_columns(ct, gc), gc)
    assert np.array_equal(to_host(g.decrypt(c, sk))[0], m[0])


@pytest.mark.gpu
def test_wire_format_moves_keys_and_ciphertexts_between_contexts(fhe, oracle):
    """keys and ciphertexts serialised by one context and loaded by a second one (a second process in production) give the
    same HMult words and the same plaintext; an object of another modulus chain is refused."""
    from fhe_b200.engine import to_device, to_host
    p, g, _ = _setup(fhe, oracle, "small")
    n, t, L, K = p["n"], p["t"], p["L"], p["K"]
    primes = list(p["primes"])
    q_chain, qp_chain, full = primes[:L], primes[:L + K], primes
    sk, pk = g.keygen(31, 32); rlk = g.relinkey_gen(33, sk)
    rng = np.random.default_rng(910)
    m1 = rng.integers(0, t, (2, n), dtype=np.uint64); m2 = rng.integers(0, t, (2, n), dtype=np.uint64)
    ca = g.encrypt(40, to_device(m1), pk); cb = g.encrypt(50, to_device(m2), pk)
    want = to_host(g.multiply(ca, cb, rlk))
    blobs = {"sk": fhe.wire_pack("secret_key", sk, full, ntt_form=True),
             "rlk": fhe.wire_pack("switch_key", rlk, qp_chain, ntt_form=True),
             "ca": fhe.wire_pack("ciphertext", ca, q_chain), "cb": fhe.wire_pack("ciphertext", cb, q_chain)}
    g2 = fhe.BfvContext(n, L, p["R"], K, p["dnum"], t, p["primes"], p["sigma"], p["hamming_weight"])
    hdr, w = fhe.wire_unpack(blobs["rlk"], qp_chain)
    assert hdr["kind"] == "switch_key" and hdr["polys"] == 2 * p["dnum"] and hdr["limbs"] == L + K
    rlk2 = to_device(w.reshape(tuple(rlk.shape)))
    sk2 = to_device(fhe.wire_unpack(blobs["sk"], full)[1].reshape(tuple(sk.shape)))
    ca2 = to_device(fhe.wire_unpack(blobs["ca"], q_chain)[1].reshape(tuple(ca.shape)))
    cb2 = to_device(fhe.wire_unpack(blobs["cb"], q_chain)[1].reshape(tuple(cb.shape)))
    out2 = g2.multiply(ca2, cb2, rlk2)
    assert np.array_equal(to_host(out2), want)
    dec = to_host(g2.decrypt(out2, sk2))
    for b in range(2):
        assert np.array_equal(dec[b], oracle.schoolbook_negacyclic(m1[b], m2[b], t))
    with pytest.raises(fhe.FheB200Error):
        fhe.wire_unpack(blobs["ca"], primes[1:L + 1])


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["small", "c2"])
def test_multiply_and_relinearize_separately_vs_oracle(fhe, oracle, preset):
    """FHEContext::multiply (three components) and ::relinearize as separate calls (src/fhe.cu:198-235): each equals the oracle word
    for word, their composition equals the fused entry point, a sum of products is relinearised once, and the squaring path
    (same operand pointer) gives the words of the general path."""
    from fhe_b200.engine import to_device, to_host
    p, g, o = _setup(fhe, oracle, preset)
    n, t, L = p["n"], p["t"], p["L"]
    sk, pk = g.keygen(61, 62); rlk = g.relinkey_gen(63, sk)
    _, osk = o.secret_keygen(61); orlk = o.relin_keygen(63, osk)
    rng = np.random.default_rng(1200)
    m = rng.integers(0, t, (4, n), dtype=np.uint64)
    ct = g.encrypt(900, to_device(m), pk)
    a, b = ct[0:2].contiguous(), ct[2:4].contiguous()
    ha, hb = to_host(a), to_host(b)
    prod3 = g.multiply_no_relin(a, b)
    assert tuple(prod3.shape) == (2, 3, L, n)
    h3 = to_host(prod3)
    for i in range(2):
        assert np.array_equal(h3[i], o.multiply(ha[i], hb[i])), i
    rel = g.relinearize(prod3, rlk)
    fused = g.multiply(a, b, rlk)
    assert np.array_equal(to_host(rel), to_host(fused))
    for i in range(2):
        assert np.array_equal(to_host(rel)[i], o.relinearize(h3[i], orlk)), i
    # lazy relinearisation: a0*b0 + a1*b1 with one key switch
    s3 = g.add3(prod3[0:1].contiguous(), prod3[1:2].contiguous())
    mods = np.array(p["primes"][:L], dtype=np.uint64)[None, :, None]
    exp3 = (h3[0] + h3[1]) % mods
    assert np.array_equal(to_host(s3)[0], exp3)
    lazy = g.relinearize(s3, rlk)
    assert np.array_equal(to_host(lazy)[0], o.relinearize(exp3, orlk))
    want = (oracle.schoolbook_negacyclic(m[0], m[2], t) + oracle.schoolbook_negacyclic(m[1], m[3], t)) % np.uint64(t)
    assert np.array_equal(to_host(g.decrypt(lazy, sk))[0], want)
    # squaring: the same tensor as both operands vs a copy of it
    sq = g.multiply(a, a, rlk)
    gen = g.multiply(a, a.clone(), rlk)
    assert np.array_equal(to_host(sq), to_host(gen))
    assert np.array_equal(to_host(sq)[0], o.multiply_relin(ha[0], ha[0], orlk))
    assert np.array_equal(to_host(g.multiply_no_relin(a, a)), to_host(g.multiply_no_relin(a, a.clone())))
    assert np.array_equal(to_host(g.decrypt(sq, sk))[1], oracle.schoolbook_negacyclic(m[1], m[1], t))
    with pytest.raises(fhe.FheB200Error):
        g.relinearize(prod3, rlk, out=prod3.view(-1)[: 2 * 2 * L * n].view(2, 2, L, n))


@pytest.mark.gpu
def test_noise_budget_matches_exact_big_integer_value(fhe, oracle):
    """fhe_b200_bfv_noise_budget against the definition evaluated with Python big integers: per limb t (c0 + c1 s) mod q_i from
    the oracle, CRT, centred, infinity norm.  Floating point only in the final log2: tolerance 1e-6 bits."""
    import math
    from math import prod
    from bigint_ref import crt
    from fhe_b200.engine import to_device, to_host
    p, g, o = _setup(fhe, oracle, "small")
    n, t, L = p["n"], p["t"], p["L"]
    qs = [int(q) for q in p["primes"][:L]]
    Q = prod(qs)
    sk, pk = g.keygen(71, 72); rlk = g.relinkey_gen(73, sk)
    s_small, _ = o.secret_keygen(71)
    rng = np.random.default_rng(1300)
    m = rng.integers(0, t, (2, n), dtype=np.uint64)
    ct = g.encrypt(950, to_device(m), pk)
    prod_ct = g.multiply(ct[0:1].contiguous(), ct[1:2].contiguous(), rlk)
    zero = torch.zeros_like(ct[0:1])                                   # (0, 0): the noise is exactly zero
    cts = torch.cat([ct, prod_ct, zero]).contiguous()
    got = g.noise_budget(cts, sk)

    def exact(h):
        worst = 0
        res = []
        for i, q in enumerate(qs):
            s_q = np.array([int(v) % q for v in s_small], dtype=np.uint64)
            v = (oracle.negacyclic_mul_ntt(h[1, i], s_q, q).astype(object) + h[0, i].astype(object)) % q
            res.append([(int(x) * t) % q for x in v])
        for k in range(n):
            w = crt([res[i][k] for i in range(L)], qs)
            worst = max(worst, min(w, Q - w))
        return math.log2(Q) - math.log2(worst) - 1 if worst else math.log2(Q) - 1

    h = to_host(cts)
    want = [exact(h[i]) for i in range(4)]
    assert np.allclose(got, want, rtol=0, atol=1e-6), (got, want)
    assert want[2] < min(want[0], want[1]) - 10 and want[2] > 0          # a product costs budget but leaves some
    assert got[3] == pytest.approx(math.log2(Q) - 1, abs=1e-9)


@pytest.mark.gpu
def test_mod_switch_to_level_vs_oracle(fhe, oracle):
    from fhe_b200.engine import to_device, to_host
    p, g, o = _setup(fhe, oracle, "small")
    n, t, L = p["n"], p["t"], p["L"]
    sk, pk = g.keygen(81, 82)
    m = np.random.default_rng(83).integers(0, t, (3, n), dtype=np.uint64)
    ct = g.encrypt(84, to_device(m), pk)
    h = to_host(ct)
    for drop in (0, 1, 2, 3):
        low = to_host(g.mod_switch_to_level(ct, drop))
        assert low.shape == (3, 2, L - drop, n)
        for b in range(3):
            assert np.array_equal(low[b], o.mod_switch_to_level(h[b], drop)), (drop, b)
    with pytest.raises(fhe.FheB200Error):
        g.mod_switch_to_level(ct, L)


@pytest.mark.parametrize("hw", [64, 0])
def test_keyed_chacha20_generator_vs_oracle(fhe, oracle, hw):
    """fhe_b200_bfv_set_rng_key: with a 256-bit key every sampler (host Fisher-Yates secret, device ternary / Gaussian / uniform,
    the errors regenerated inside encrypt) draws from ChaCha20; keys and ciphertexts equal the oracle's keyed mode word for word,
    differ from the splitmix mode, and switching the key off restores the reproducible generator."""
    from fhe_b200.engine import to_device, to_host
    p, g, o = _setup(fhe, oracle, "small", hw)
    n, t = p["n"], p["t"]
    key = bytes((7 * i + 3) & 0xFF for i in range(32))
    sk0, pk0 = g.keygen(11, 12)
    m = np.random.default_rng(5).integers(0, t, (3, n), dtype=np.uint64)
    try:
        g.set_rng_key(key); oracle.set_rng_key(key)
        sk, pk = g.keygen(11, 12)
        _, osk = o.secret_keygen(11); opk = o.public_keygen(12, osk)
        assert np.array_equal(to_host(sk), osk) and np.array_equal(to_host(pk), opk)
        assert not np.array_equal(to_host(sk), to_host(sk0)) and not np.array_equal(to_host(pk), to_host(pk0))
        rlk = g.relinkey_gen(13, sk)
        assert np.array_equal(to_host(rlk), o.relin_keygen(13, osk))
        ct = g.encrypt(500, to_device(m), pk)
        h = to_host(ct)
        for b in range(3):
            assert np.array_equal(h[b], o.encrypt(500, m[b], opk, item=b)), b
        assert np.array_equal(to_host(g.decrypt(ct, sk)), m)
        prod = g.multiply(ct[0:1].contiguous(), ct[1:2].contiguous(), rlk)
        assert np.array_equal(to_host(g.decrypt(prod, sk))[0], oracle.schoolbook_negacyclic(m[0], m[1], t))
    finally:
        g.set_rng_key(None); oracle.set_rng_key(None)
    sk1, pk1 = g.keygen(11, 12)
    assert torch.equal(sk1, sk0) and torch.equal(pk1, pk0)
