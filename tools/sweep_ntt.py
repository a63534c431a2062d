"""BASELINE.json config 5 on one GPU: batched RNS-NTT throughput over N = 2^12..2^17 and batch sizes (ciphertexts of 2 polynomials,
L(N) limbs as SURVEY 8d suggests).  One JSON line per point: forward+inverse limb-transforms/s, CUDA events, round trip checked.
    python tools/sweep_ntt.py [--max-bytes 8e9]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fhe_b200
from fhe_b200.params import prime_chain

ap = argparse.ArgumentParser(); ap.add_argument("--max-bytes", type=float, default=8e9); a = ap.parse_args()
LIMBS = {12: 2, 13: 4, 14: 8, 15: 16, 16: 24, 17: 32}
chain = prime_chain(32)
for logn, L in LIMBS.items():
    n = 1 << logn
    plan = fhe_b200.Plan(n, chain[:L])
    for batch in (1, 4, 16, 64, 256, 1024):
        polys = 2 * batch
        if polys * L * n * 8 > a.max_bytes:
            continue
        g = torch.Generator(device="cuda"); g.manual_seed(logn * 1000 + batch)
        x = torch.empty((polys, L, n), dtype=torch.int64, device="cuda")
        for l, q in enumerate(chain[:L]):
            x[:, l, :] = torch.randint(0, q, (polys, n), generator=g, device="cuda", dtype=torch.int64)
        ref = x.clone()
        for _ in range(3):
            plan.forward(x); plan.inverse(x)
        torch.cuda.synchronize()
        units = 2 * polys * L
        reps = max(3, min(200, int(2e5 / units)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            plan.forward(x); plan.inverse(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        ok = bool(torch.equal(x, ref))
        mb = polys * L * n * 8 / 2**20
        print(json.dumps({"logn": logn, "limbs": L, "ciphertexts": batch, "polys": polys, "working_set_MiB": round(mb, 1),
                          "ms_fwd_inv": round(ms, 4), "limb_transforms_per_s": round(units / (ms / 1e3)),
                          "butterflies_per_s_G": round(units / (ms / 1e3) * (n // 2) * logn / 1e9, 1), "roundtrip_bit_exact": ok}), flush=True)
        del x, ref
    del plan
    torch.cuda.empty_cache()
