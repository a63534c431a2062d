"""Committed golden digests (tests/golden/golden.json, written by tests/golden/make_golden.py): the oracle must reproduce them
(CPU) and the CUDA path must produce the same bytes (GPU).  The reference ships no fixture files; these pin the restatement."""
import hashlib
import importlib.util
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def test_oracle_reproduces_golden_digests(oracle):
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    got = mod.cases()
    assert set(got) == set(GOLD)
    for k in GOLD:
        assert got[k] == GOLD[k], k


@pytest.mark.gpu
def test_cuda_path_matches_golden_digests(oracle):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import fhe_b200
    from fhe_b200.engine import to_device, to_host
    from fhe_b200.params import bfv_preset
    chain = oracle.prime_chain(49)
    for logn in (10, 12, 13, 16):
        n = 1 << logn
        for li in (0, 31):
            q = chain[li]
            x = np.random.default_rng(0x5EED0000 + logn * 100 + li).integers(0, q, n, dtype=np.uint64)
            g = GOLD[f"ntt_fwd/logn{logn}/P{li}"]
            assert _digest(x) == g["input"]
            d = to_device(x.reshape(1, 1, n)); fhe_b200.Plan(n, [q]).forward(d)
            assert _digest(to_host(d)) == g["output"], (logn, li)
    for n, q in ((1024, 12289), (2048, 40961)):
        x = np.arange(1, n + 1, dtype=np.uint64) % np.uint64(q)
        d = to_device(x.reshape(1, 1, n)); fhe_b200.Plan(n, [q]).forward(d)
        assert _digest(to_host(d)) == GOLD[f"ntt_fwd/ref_n{n}_q{q}"]["output"]
    src, dst = chain[:24], chain[24:49]
    rng = np.random.default_rng(0x5EED0100)
    z = np.stack([rng.integers(0, m, 256, dtype=np.uint64) for m in src])
    assert _digest(z) == GOLD["conv/24to25"]["input"]
    assert _digest(to_host(fhe_b200.LinComb.conv(src, dst).apply(to_device(z[None])))[0]) == GOLD["conv/24to25"]["output"]
    zp = np.stack([rng.integers(0, m, 256, dtype=np.uint64) for m in dst])
    got = to_host(fhe_b200.LinComb.scale(src, dst, 65537, dst, True).apply(to_device(z[None]), to_device(zp[None])))[0]
    assert _digest(got) == GOLD["scale/24to25_t65537"]["output"]
    p = bfv_preset("c2")
    g = fhe_b200.BfvContext(p["n"], p["L"], p["R"], p["K"], p["dnum"], p["t"], p["primes"], p["sigma"], p["hamming_weight"])
    sk, pk = g.keygen(11, 12); rlk = g.relinkey_gen(13, sk); gk = g.galoiskey_gen(14, 3, sk)
    m1 = np.random.default_rng(15).integers(0, p["t"], p["n"], dtype=np.uint64)
    m2 = np.random.default_rng(16).integers(0, p["t"], p["n"], dtype=np.uint64)
    c1 = g.encrypt(17, to_device(m1[None]), pk); c2 = g.encrypt(18, to_device(m2[None]), pk)
    assert _digest(to_host(sk)) == GOLD["bfv_c2/sk"]["output"]
    assert _digest(to_host(pk)) == GOLD["bfv_c2/pk"]["output"]
    assert _digest(to_host(rlk)) == GOLD["bfv_c2/rlk"]["output"]
    assert _digest(to_host(gk)) == GOLD["bfv_c2/gk3"]["output"]
    assert _digest(to_host(c1)[0]) == GOLD["bfv_c2/encrypt"]["output"]
    assert _digest(to_host(g.multiply(c1, c2, rlk))[0]) == GOLD["bfv_c2/multiply_relin"]["output"]
    prod3 = g.multiply_no_relin(c1, c2)
    assert _digest(to_host(prod3)[0]) == GOLD["bfv_c2/multiply_3_components"]["output"]
    lazy = g.relinearize(g.add3(prod3, g.multiply_no_relin(c1, c1)), rlk)
    assert _digest(to_host(lazy)[0]) == GOLD["bfv_c2/relinearize_sum_of_products"]["output"]
    assert _digest(to_host(g.multiply(c1, c1, rlk))[0]) == GOLD["bfv_c2/square"]["output"]
    assert _digest(to_host(g.apply_galois(c1, 3, gk))[0]) == GOLD["bfv_c2/rotate3"]["output"]
    assert _digest(to_host(g.mod_switch_to_next(c1))[0]) == GOLD["bfv_c2/mod_switch_to_next"]["output"]
