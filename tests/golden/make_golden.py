"""Regenerates tests/golden/golden.json: SHA-256 digests of oracle outputs on seeded inputs (plus the inputs' own digests), so that
(a) the oracle cannot drift silently and (b) the CUDA path can be checked against committed values on a box that has no reference
checkout.  The reference ships no fixture files; its known answers are asserted directly in tests/test_oracle_*.py.
    python tests/golden/make_golden.py"""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def cases():
    chain = oracle.prime_chain(49)
    out = {}
    for logn in (10, 12, 13, 16):
        n = 1 << logn
        for li in (0, 31):
            q = chain[li]
            x = np.random.default_rng(0x5EED0000 + logn * 100 + li).integers(0, q, n, dtype=np.uint64)
            y = oracle.ntt_forward(x, q)
            out[f"ntt_fwd/logn{logn}/P{li}"] = {"input": digest(x), "output": digest(y)}
    # the reference's own test moduli and pattern x_i = i + 1 (tests/test_fhe.cu:68-78, 129-139)
    for n, q in ((1024, 12289), (2048, 40961)):
        x = (np.arange(1, n + 1, dtype=np.uint64)) % np.uint64(q)
        out[f"ntt_fwd/ref_n{n}_q{q}"] = {"input": digest(x), "output": digest(oracle.ntt_forward(x, q))}
    # exact base conversion Q(24) -> R(25) and scale-and-round, 256 coefficients
    src, dst = chain[:24], chain[24:49]
    rng = np.random.default_rng(0x5EED0100)
    z = np.stack([rng.integers(0, m, 256, dtype=np.uint64) for m in src])
    out["conv/24to25"] = {"input": digest(z), "output": digest(oracle.LinComb.conv(src, dst).apply(z))}
    zp = np.stack([rng.integers(0, m, 256, dtype=np.uint64) for m in dst])
    out["scale/24to25_t65537"] = {"input": digest(np.concatenate([z, zp])),
                                  "output": digest(oracle.LinComb.scale(src, dst, 65537, dst, True).apply(z, extra=zp))}
    # BFV at BASELINE config 2 (N=4096, L=2): keys, ciphertexts, product, rotation
    from importlib import import_module
    sys.path.insert(0, os.path.join(ROOT))
    params = import_module("fhe_b200.params").bfv_preset("c2")
    o = oracle.Bfv(params["n"], params["L"], params["R"], params["K"], params["dnum"], params["t"], params["primes"],
                   sigma=params["sigma"], hw=params["hamming_weight"])
    _, sk = o.secret_keygen(11); pk = o.public_keygen(12, sk); rlk = o.relin_keygen(13, sk); gk = o.galois_keygen(14, 3, sk)
    m1 = np.random.default_rng(15).integers(0, params["t"], params["n"], dtype=np.uint64)
    m2 = np.random.default_rng(16).integers(0, params["t"], params["n"], dtype=np.uint64)
    c1 = o.encrypt(17, m1, pk); c2 = o.encrypt(18, m2, pk)
    out["bfv_c2/sk"] = {"output": digest(sk)}; out["bfv_c2/pk"] = {"output": digest(pk)}; out["bfv_c2/rlk"] = {"output": digest(rlk)}
    out["bfv_c2/gk3"] = {"output": digest(gk)}
    out["bfv_c2/encrypt"] = {"input": digest(m1), "output": digest(c1)}
    out["bfv_c2/multiply_relin"] = {"output": digest(o.multiply_relin(c1, c2, rlk))}
    prod3 = o.multiply(c1, c2)
    out["bfv_c2/multiply_3_components"] = {"output": digest(prod3)}
    mods = np.array(params["primes"][:params["L"]], dtype=np.uint64)[None, :, None]
    out["bfv_c2/relinearize_sum_of_products"] = {"output": digest(o.relinearize((prod3 + o.multiply(c1, c1)) % mods, rlk))}
    out["bfv_c2/square"] = {"output": digest(o.multiply_relin(c1, c1, rlk))}
    out["bfv_c2/rotate3"] = {"output": digest(o.apply_galois(c1, 3, gk))}
    out["bfv_c2/mod_switch_to_next"] = {"output": digest(o.mod_switch_to_next(c1))}
    return out


def mix64(z):
    z = z.astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def config3_input(chain, polys=64, limbs=32, n=1 << 16):
    """BASELINE config 3 input uint64[polys][limbs][n]: x = mix64(0x5EED0003 + flat index) mod q_limb (host formula, so that the GPU
    test and the oracle transform the very same gigabyte)"""
    x = np.empty((polys, limbs, n), dtype=np.uint64)
    j = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for b in range(polys):
            for l in range(limbs):
                x[b, l] = mix64(np.uint64(0x5EED0003) + np.uint64((b * limbs + l) * n) + j) % np.uint64(chain[l])
    return x


def slow_cases():
    """full-size BASELINE configs 3 and 4: the oracle needs a minute and a few GiB here, so these digests are regenerated only on
    request (python tests/golden/make_golden.py --slow) and checked by the GPU tests, not by the CPU suite."""
    out = {}
    chain = oracle.prime_chain(49)
    x = config3_input(chain)
    eng = oracle.RnsNtt(1 << 16, chain[:32])
    flat = x.reshape(-1).copy()
    eng.run_inplace(flat, 64, False, 0)
    out["config3/forward_all_64x32"] = {"input": digest(x), "output": digest(flat)}
    from importlib import import_module
    p = import_module("fhe_b200.params").bfv_preset("c4")
    o = oracle.Bfv(p["n"], p["L"], p["R"], p["K"], p["dnum"], p["t"], p["primes"], sigma=p["sigma"], hw=p["hamming_weight"])
    _, sk = o.secret_keygen(31); pk = o.public_keygen(32, sk); rlk = o.relin_keygen(33, sk)
    m1 = np.random.default_rng(34).integers(0, p["t"], p["n"], dtype=np.uint64)
    m2 = np.random.default_rng(35).integers(0, p["t"], p["n"], dtype=np.uint64)
    c1 = o.encrypt(36, m1, pk); c2 = o.encrypt(37, m2, pk)
    prod, scaled = o.multiply_relin(c1, c2, rlk, want_scaled=True)
    out["bfv_c4/sk"] = {"output": digest(sk)}; out["bfv_c4/pk"] = {"output": digest(pk)}; out["bfv_c4/rlk"] = {"output": digest(rlk)}
    out["bfv_c4/encrypt"] = {"input": digest(m1), "output": digest(c1)}
    out["bfv_c4/scaled_tensor"] = {"output": digest(scaled)}
    out["bfv_c4/multiply_relin"] = {"output": digest(prod)}
    out["bfv_c4/square"] = {"output": digest(o.multiply_relin(c1, c1, rlk))}
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    p = os.path.join(here, "golden.json")
    json.dump(cases(), open(p, "w"), indent=1, sort_keys=True)
    print("wrote", p)
    if "--slow" in sys.argv:
        p = os.path.join(here, "golden_slow.json")
        json.dump(slow_cases(), open(p, "w"), indent=1, sort_keys=True)
        print("wrote", p)
