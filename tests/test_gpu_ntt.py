"""GPU parity tests (run on the B200 with -m gpu): the CUDA path through the C ABI vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def fhe():
    import fhe_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    fhe_b200.load_library()     # must exist on a GPU box: no fallback
    return fhe_b200


def _rand(rng, mods, n, batch):
    return np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods]) for _ in range(batch)])


@pytest.mark.parametrize("logn,limbs,batch", [(9, 2, 3), (10, 1, 1), (11, 3, 2), (12, 2, 5), (13, 3, 2), (14, 2, 3),
                                              (15, 2, 2), (16, 3, 2), (17, 2, 1)])
def test_forward_inverse_vs_oracle(fhe, oracle, chain, logn, limbs, batch):
    from fhe_b200.engine import to_device, to_host
    n = 1 << logn
    mods = chain[5:5 + limbs]
    rng = np.random.default_rng(0x5EED0000 + logn)
    x = _rand(rng, mods, n, batch)
    x[0, 0, :] = mods[0] - 1                      # worst case for the lazy bounds
    plan = fhe.Plan(n, mods)
    d = to_device(x)
    plan.forward(d)
    y = to_host(d)
    for b in range(batch):
        for l, q in enumerate(mods):
            assert np.array_equal(y[b, l], oracle.ntt_forward(x[b, l], q)), (b, l)
    out = torch.empty_like(d)
    plan.inverse(d, out=out)                      # out of place
    assert np.array_equal(to_host(out), x)
    assert np.array_equal(to_host(d), y)          # input untouched


@pytest.mark.parametrize("env", [{"FHE_B200_NTT_BAL": "0"}, {"FHE_B200_NO_NEAR60": "1"},
                                 {"FHE_B200_NTT_BAL": "0", "FHE_B200_NO_NEAR60": "1"}, {"FHE_B200_NTT_CHUNK_MB": "1"},
                                 {"FHE_B200_NTT_BAL": "0", "FHE_B200_NTT_CHUNK_MB": "1"}])
def test_alternative_code_paths(fhe, oracle, chain, env, monkeypatch):
    """the row+tile two-pass strategy (FHE_B200_NTT_BAL=0), chunked launches and the generic (non near-2^60) reduction are selected
    PER PLAN from the environment when the plan is created (so setting them here, before Plan(), really switches the path); they
    must give the same bits as the default balanced / near-2^60 path."""
    from fhe_b200.engine import to_device, to_host
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    n, mods = 1 << 14, chain[:3]
    rng = np.random.default_rng(77)
    x = _rand(rng, mods, n, 11)                       # 11 polynomials: a ragged last group of the fused scheduler's 8
    plan = fhe.Plan(n, mods)
    d = to_device(x); plan.forward(d); y = to_host(d)
    for b in (0, 7, 10):
        for l, q in enumerate(mods):
            assert np.array_equal(y[b, l], oracle.ntt_forward(x[b, l], q))
    plan.inverse(d)
    assert np.array_equal(to_host(d), x)


@pytest.mark.parametrize("logn", [13, 14, 15, 16])
def test_balanced_passes_ragged_batches_and_limb_ranges(fhe, oracle, chain, logn):
    """ntt_bal: item runs of pass A and polynomial groups of pass B with batches that do not divide evenly, 61-bit and
    far-from-2^60 moduli (generic reduction), limb sub-ranges, in place and out of place."""
    from fhe_b200.engine import to_device, to_host
    n = 1 << logn
    rng = np.random.default_rng(900 + logn)
    p61 = oracle.prime_chain(2, bits=61)
    small = [p for p in range(2 * n + 1, 1 << 24, 2 * n) if oracle.is_prime(p)][:2]
    for mods, batches in ((chain[:5], (1, 3, 7, 20)), (p61, (2,)), (small + chain[:1], (5,))):
        plan = fhe.Plan(n, mods)
        for batch in batches:
            x = _rand(rng, mods, n, batch)
            d = to_device(x); out = torch.empty_like(d)
            plan.forward(d, out=out)
            y = to_host(out)
            for b in {0, batch - 1}:
                for l, q in enumerate(mods):
                    assert np.array_equal(y[b, l], oracle.ntt_forward(x[b, l], q)), (batch, b, l)
            plan.inverse(out)                         # in place
            assert np.array_equal(to_host(out), x)
    # limb sub-range: buffer limbs map to plan limbs [2, 4)
    mods = chain[:5]
    plan = fhe.Plan(n, mods)
    x = _rand(rng, mods[2:4], n, 3)
    d = to_device(x)
    plan.forward(d, limb_begin=2)
    y = to_host(d)
    for l, q in enumerate(mods[2:4]):
        assert np.array_equal(y[1, l], oracle.ntt_forward(x[1, l], q))
    plan.inverse(d, limb_begin=2)
    assert np.array_equal(to_host(d), x)


def test_tables_match_oracle(fhe, oracle, chain):
    n = 2048
    plan = fhe.Plan(n, chain[:2])
    for l, q in enumerate(chain[:2]):
        fw, fws, iv, ivs = plan.tables(l)
        of, oi = oracle.ntt_tables(q, n)
        assert np.array_equal(fw, of) and np.array_equal(iv, oi)
        assert [int(v) for v in fws[:64]] == [(int(w) << 64) // q for w in of[:64]]
        assert [int(v) for v in ivs[:64]] == [(int(w) << 64) // q for w in oi[:64]]


def test_headroom_8_path_61_bit_primes(fhe, oracle):
    from fhe_b200.engine import to_device, to_host
    mods = oracle.prime_chain(2, bits=61)
    for logn in (12, 16):
        n = 1 << logn
        rng = np.random.default_rng(61 + logn)
        x = _rand(rng, mods, n, 2); x[1, 1, :] = mods[1] - 1
        plan = fhe.Plan(n, mods)
        d = to_device(x); plan.forward(d); y = to_host(d)
        for b in range(2):
            for l, q in enumerate(mods):
                assert np.array_equal(y[b, l], oracle.ntt_forward(x[b, l], q))
        plan.inverse(d)
        assert np.array_equal(to_host(d), x)


def test_reference_test_vectors(fhe, oracle):
    """the reference's own NTT checks: N=1024,q=12289,x_i=i+1 round trip (tests/test_fhe.cu:68-116) and
    N=2048,q=40961 product (tests/test_fhe.cu:126-167), here actually compared."""
    from fhe_b200.engine import to_device, to_host
    n, q = 1024, 12289
    x = np.arange(1, n + 1, dtype=np.uint64)
    eng = fhe.NTTEngine(n, q)
    d = to_device(x)
    eng.forward(d)
    assert np.array_equal(to_host(d), oracle.ntt_forward(x, q))
    eng.inverse(d)
    assert np.array_equal(to_host(d), x)
    n, q = 2048, 40961
    rng = np.random.default_rng(7)
    a = rng.integers(0, 100, n, dtype=np.uint64); b = rng.integers(0, 100, n, dtype=np.uint64)
    eng = fhe.NTTEngine(n, q)
    da, db = to_device(a), to_device(b); dc = torch.empty_like(da)
    eng.multiply(dc, da, db)
    assert np.array_equal(to_host(dc), oracle.schoolbook_negacyclic(a, b, q))
    assert np.array_equal(to_host(da), a) and np.array_equal(to_host(db), b)


def test_config1_negacyclic_product_vs_schoolbook(fhe, oracle, chain):
    """BASELINE.json config 1."""
    from fhe_b200.engine import to_device, to_host
    n, q = 4096, chain[0]
    rng = np.random.default_rng(0x5EED0001)
    a = rng.integers(0, q, n, dtype=np.uint64); b = rng.integers(0, q, n, dtype=np.uint64)
    eng = fhe.NTTEngine(n, q)
    dc = to_device(np.zeros(n, dtype=np.uint64))
    eng.multiply(dc, to_device(a), to_device(b))
    assert np.array_equal(to_host(dc), oracle.schoolbook_negacyclic(a, b, q))
    # aliasing result with an input is allowed (reference copies inputs first, src/ntt.cu:55-58)
    da = to_device(a)
    eng.multiply(da, da, to_device(b))
    assert np.array_equal(to_host(da), oracle.schoolbook_negacyclic(a, b, q))


@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("logn", [9, 10, 11, 12])
def test_negacyclic_mul_fused_and_unfused(fhe, oracle, chain, logn, fused, monkeypatch):
    """fhe_b200_negacyclic_mul at N <= 4096: the single fused kernel (two forward tile passes, pointwise product and inverse in
    shared memory) and the four-launch path give the schoolbook product bit for bit; 60- and 61-bit moduli, several limbs and
    polynomials, worst-case inputs, in place on either operand."""
    from fhe_b200.engine import to_device, to_host
    monkeypatch.setenv("FHE_B200_MUL_FUSED", fused)
    n = 1 << logn
    rng = np.random.default_rng(40 + logn)
    for mods in (chain[:3], oracle.prime_chain(2, bits=61)):
        plan = fhe.Plan(n, mods)
        a = _rand(rng, mods, n, 2); b = _rand(rng, mods, n, 2)
        a[0, :, :] = np.array(mods, dtype=np.uint64)[:, None] - np.uint64(1); b[0, 0, :] = np.uint64(mods[0] - 1)
        want = np.stack([np.stack([oracle.negacyclic_mul_ntt(a[i, l], b[i, l], q) for l, q in enumerate(mods)]) for i in range(2)])
        assert np.array_equal(want[1, 0], oracle.schoolbook_negacyclic(a[1, 0], b[1, 0], mods[0]))
        da, db = to_device(a), to_device(b)
        out = plan.negacyclic_mul(da, db)
        assert np.array_equal(to_host(out), want)
        plan.negacyclic_mul(da, db, out=da)               # in place on the first operand
        assert np.array_equal(to_host(da), want)
        da = to_device(a)
        plan.negacyclic_mul(da, db, out=db)               # ... and on the second
        assert np.array_equal(to_host(db), want)


@pytest.mark.parametrize("vec", ["zero", "qm1", "delta0", "deltaN", "iota"])
def test_edge_vectors(fhe, oracle, chain, vec):
    from fhe_b200.engine import to_device, to_host
    n, q = 8192, chain[3]
    a = np.zeros(n, dtype=np.uint64)
    if vec == "qm1": a[:] = q - 1
    if vec == "delta0": a[0] = 1
    if vec == "deltaN": a[n - 1] = 1
    if vec == "iota": a = np.arange(1, n + 1, dtype=np.uint64)
    eng = fhe.NTTEngine(n, q)
    d = to_device(a); eng.forward(d)
    assert np.array_equal(to_host(d), oracle.ntt_forward(a, q))
    eng.inverse(d)
    assert np.array_equal(to_host(d), a)


def test_sub_range_of_limbs_and_empty_batch(fhe, oracle, chain):
    from fhe_b200.engine import to_device, to_host
    n = 4096
    plan = fhe.Plan(n, chain[:6])
    rng = np.random.default_rng(11)
    x = _rand(rng, chain[2:5], n, 2)
    d = to_device(x)
    plan.forward(d, limb_begin=2, limb_count=3)
    y = to_host(d)
    for b in range(2):
        for l in range(3):
            assert np.array_equal(y[b, l], oracle.ntt_forward(x[b, l], chain[2 + l]))
    empty = torch.empty((0, 3, n), dtype=torch.int64, device="cuda")
    plan.forward(empty, limb_begin=2, limb_count=3)          # no-op, no error
    with pytest.raises(fhe.FheB200Error):
        plan.forward(d, limb_begin=5, limb_count=3)         # out of range


def test_bad_parameters_are_rejected(fhe, chain):
    with pytest.raises(fhe.FheB200Error):
        fhe.Plan(4096, [chain[0] - 2])          # not prime
    with pytest.raises(fhe.FheB200Error):
        fhe.Plan(4096, [12289])                 # prime but 12289 != 1 mod 8192
    with pytest.raises(fhe.FheB200Error):
        fhe.Plan(3000, [chain[0]])              # not a power of two
    with pytest.raises(fhe.FheB200Error):
        fhe.Plan(1 << 18, [chain[0]])           # too large


def test_elementwise_vs_oracle(fhe, oracle, chain):
    from fhe_b200.engine import to_device, to_host
    n, mods = 2048, chain[:3] + [12289 * 0 + 786433]
    plan = fhe.Plan(n, mods)
    rng = np.random.default_rng(12)
    a = _rand(rng, mods, n, 3); b = _rand(rng, mods, n, 3); c = _rand(rng, mods, n, 3)
    a[0, 0, :4] = [0, mods[0] - 1, 1, mods[0] - 1]; b[0, 0, :4] = [0, mods[0] - 1, mods[0] - 1, 1]
    da, db, dc = to_device(a), to_device(b), to_device(c)
    assert np.array_equal(to_host(plan.add(da, db)), oracle.poly_add(a, b, mods, n))
    assert np.array_equal(to_host(plan.sub(da, db)), oracle.poly_sub(a, b, mods, n))
    assert np.array_equal(to_host(plan.mul(da, db)), oracle.poly_mul(a, b, mods, n))
    assert np.array_equal(to_host(plan.mac(dc, da, db)), oracle.poly_mac(c, a, b, mods, n))
    sc = [2**59 + 12345, 3, 0, 2**63 + 5]
    assert np.array_equal(to_host(plan.mul_scalar(da, sc)), oracle.poly_mul_scalar(a, [s % m for s, m in zip(sc, mods)], mods, n))
    neg = to_host(plan.negate(da))
    assert np.array_equal(to_host(plan.add(to_device(neg), da)), np.zeros_like(a))
    assert np.array_equal(to_host(plan.add_scalar(da, [1, 1, 1, 1])), oracle.poly_add(a, np.ones_like(a), mods, n))


def test_u256_edge_and_bitrev(fhe, oracle, chain):
    from fhe_b200.engine import pack_u256, to_device, to_host, unpack_u256
    n = 1024
    rng = np.random.default_rng(13)
    v = rng.integers(0, 2**61, n, dtype=np.uint64)
    packed = pack_u256(to_device(v))
    w = to_host(packed).reshape(n, 4)
    assert np.array_equal(w[:, 0], v) and not w[:, 1:].any()
    assert np.array_equal(to_host(unpack_u256(packed)), v)
    # to_rns of full 256-bit values
    words = rng.integers(0, 2**64, (n, 4), dtype=np.uint64)
    eng = fhe.RNS_NTTEngine(n, chain[:3])
    res = to_host(eng.to_rns(to_device(words)))
    for l, q in enumerate(chain[:3]):
        exp = [(int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192) % q for r in words[:32]]
        assert [int(x) for x in res[l, :32]] == exp
    # CRT back: from_rns(to_rns(x)) == x mod Q
    Q = chain[0] * chain[1] * chain[2]
    back = to_host(eng.from_rns(to_device(res))).reshape(n, 4)
    for r, b in list(zip(words, back))[:64]:
        x = int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192
        assert int(b[0]) | int(b[1]) << 64 | int(b[2]) << 128 | int(b[3]) << 192 == x % Q
    # natural-order NTT values = definitional transform
    q = 12289
    x = rng.integers(0, q, n, dtype=np.uint64)
    plan = fhe.Plan(n, [q])
    d = to_device(x.reshape(1, 1, n)); plan.forward(d)
    nat = to_host(plan.bitrev_permute(d)).reshape(n)
    assert np.array_equal(nat, oracle.negacyclic_dft_def(x, q))


def test_host_buffer_path(fhe, oracle, chain):
    from fhe_b200.engine import pinned_empty
    n, mods, batch = 1 << 14, chain[:4], 80        # 80 polys x 4 limbs x 128 KiB = 40 MiB -> several pipeline chunks
    plan = fhe.Plan(n, mods)
    rng = np.random.default_rng(14)
    h = pinned_empty((batch, len(mods), n))
    for l, q in enumerate(mods):
        h[:, l, :] = rng.integers(0, q, (batch, n), dtype=np.uint64)
    x = h.copy()
    plan.ntt_host(h, 0)
    for b in (0, 37, 79):
        for l, q in enumerate(mods):
            assert np.array_equal(h[b, l], oracle.ntt_forward(x[b, l], q))
    plan.ntt_host(h, 1)
    assert np.array_equal(h, x)
    plan.ntt_host(h, 2)
    assert np.array_equal(h, x)


@pytest.mark.slow
def test_config3_full_size_properties(fhe, oracle, chain):
    """BASELINE.json config 3 at full size: [64][32][2^16] (1 GiB).  Size-independent properties: round trip,
    linearity of the forward transform, plus oracle spot checks on sampled limbs."""
    n, limbs, batch = 1 << 16, 32, 64
    mods = chain[:limbs]
    plan = fhe.Plan(n, mods)
    g = torch.Generator(device="cuda"); g.manual_seed(0x5EED0003)
    x = torch.empty((batch, limbs, n), dtype=torch.int64, device="cuda")
    for l, q in enumerate(mods):
        x[:, l, :] = torch.randint(0, q, (batch, n), generator=g, device="cuda", dtype=torch.int64)
    ref = x.clone()
    plan.forward(x)
    for (b, l) in [(0, 0), (63, 31), (17, 5)]:
        got = x[b, l].cpu().numpy().view(np.uint64)
        assert np.array_equal(got, oracle.ntt_forward(ref[b, l].cpu().numpy().view(np.uint64), mods[l]))
    # linearity: NTT(x0 + x1) == NTT(x0) + NTT(x1) on one polynomial
    s = plan.add(ref[0:1].contiguous(), ref[1:2].contiguous())
    plan.forward(s)
    assert torch.equal(s, plan.add(x[0:1].contiguous(), x[1:2].contiguous()))
    plan.inverse(x)
    assert torch.equal(x, ref)


@pytest.mark.gpu
def test_config3_whole_forward_output_digest(fhe, oracle, chain):
    """BASELINE.json config 3: SHA-256 of the WHOLE forward output (2048 limb-polynomials, 1 GiB) equals the digest the oracle
    produced once (tests/golden/make_golden.py --slow, committed in golden_slow.json) -- every word pinned, not a sample."""
    import hashlib, importlib.util, json, os
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", "golden_slow.json")))["config3/forward_all_64x32"]
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    x = mg.config3_input(chain)
    assert mg.digest(x) == gold["input"]
    plan = fhe.Plan(1 << 16, chain[:32])
    d = torch.from_numpy(x.view(np.int64)).cuda()
    plan.forward(d)
    y = d.cpu().numpy().view(np.uint64)
    assert hashlib.sha256(y.tobytes()).hexdigest() == gold["output"]
    plan.inverse(d)
    assert np.array_equal(d.cpu().numpy().view(np.uint64), x)


@pytest.mark.gpu
def test_two_devices_in_one_process(fhe, oracle):
    """plans, conversions and BFV contexts on two GPUs of one process (the C ABI takes a device ordinal): function attributes
    such as the dynamic shared-memory limit are per device.  Skipped on a single-GPU box."""
    import torch
    from fhe_b200.engine import to_device, to_host
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    chain = oracle.prime_chain(49)
    rng = np.random.default_rng(4242)
    for dev in (0, 1, 0):
        with torch.cuda.device(dev):
            for n in (1 << 12, 1 << 13, 1 << 16):
                plan = fhe.Plan(n, chain[:2], device=dev)
                x = np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in chain[:2]])[None]
                dx = to_device(x, f"cuda:{dev}")
                plan.forward(dx)
                torch.cuda.synchronize()
                assert np.array_equal(to_host(dx)[0, 1], oracle.ntt_forward(x[0, 1], chain[1])), (dev, n)
                plan.inverse(dx)
                torch.cuda.synchronize()
                assert np.array_equal(to_host(dx), x), (dev, n)
            src, dst = chain[:24], chain[24:]
            z = np.stack([rng.integers(0, m, 256, dtype=np.uint64) for m in src])[None]
            got = to_host(fhe.LinComb.conv(src, dst, device=dev).apply(to_device(z, f"cuda:{dev}")))[0]
            assert np.array_equal(got, oracle.LinComb.conv(src, dst).apply(z[0])), dev
    # a BFV context on the second device while the current device is the first: the entry points switch to the object's
    # device and restore the caller's
    from fhe_b200.params import bfv_preset
    p = bfv_preset("small")
    torch.cuda.set_device(0)
    g = fhe.BfvContext(p["n"], p["L"], p["R"], p["K"], p["dnum"], p["t"], p["primes"], p["sigma"], p["hamming_weight"], device=1)
    assert torch.cuda.current_device() == 0
    with torch.cuda.device(1):
        sk, pk = g.keygen(1, 2); rlk = g.relinkey_gen(3, sk)
        m = rng.integers(0, p["t"], (2, p["n"]), dtype=np.uint64)
        ct = g.encrypt(5, to_device(m, "cuda:1"), pk)
        out = g.multiply(ct[0:1].contiguous(), ct[1:2].contiguous(), rlk)
        dec = to_host(g.decrypt(out, sk))[0]
    assert np.array_equal(dec, oracle.schoolbook_negacyclic(m[0], m[1], p["t"]))
    assert torch.cuda.current_device() == 0


def test_measured_integer_peaks_and_butterfly_loop(fhe, chain):
    """The roofline denominators bench.py measures in its own run (csrc/peaks.cu): IMAD and IMAD.WIDE rates, and the engine's own lazy
    butterfly in a register-only loop, which cannot beat the mix of the two it is made of (5 IMAD.WIDE + 4 IMAD per butterfly)."""
    import ctypes as C
    lib = fhe.load_library()
    lo, wide, loop = C.c_double(), C.c_double(), C.c_double()
    fhe.check(lib.fhe_b200_measure_int_peaks(0, C.byref(lo), C.byref(wide), 3))
    lib.fhe_b200_measure_butterfly_loop.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_double), C.c_int]
    fhe.check(lib.fhe_b200_measure_butterfly_loop(0, int(chain[0]), int(chain[0]) // 3, C.byref(loop), 3))
    assert lo.value > 5e12 and wide.value > 2e12 and lo.value > wide.value
    mix = 1.0 / (5.0 / wide.value + 4.0 / lo.value)                      # butterflies/s if the multiplier were all there is
    assert 0.5 * mix < loop.value <= 1.02 * mix
    with pytest.raises(fhe.FheB200Error):
        fhe.check(lib.fhe_b200_measure_butterfly_loop(0, 12289, 3, C.byref(loop), 1))   # not a chain prime
