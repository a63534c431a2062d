"""Experiment: do two independent BFV contexts on two streams overlap (tensor-core conversions of one with the NTT passes of the
other)?  Aggregate HMult+relinearize ops/s of 2 contexts x batch B on two streams vs 1 context x batch 2B on one stream."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fhe_b200
from fhe_b200.engine import to_device
from fhe_b200.params import bfv_preset
p = bfv_preset("c4"); n, L, t = p["n"], p["L"], p["t"]
def mk(B, seed):
    g = fhe_b200.BfvContext(n, L, p["R"], p["K"], p["dnum"], t, p["primes"], p["sigma"], p["hamming_weight"])
    sk, pk = g.keygen(seed, seed + 1); rlk = g.relinkey_gen(seed + 2, sk)
    rng = np.random.default_rng(seed)
    ca = g.encrypt(5, to_device(rng.integers(0, t, (B, n), dtype=np.uint64)), pk)
    cb = g.encrypt(6, to_device(rng.integers(0, t, (B, n), dtype=np.uint64)), pk)
    out = torch.empty_like(ca)
    g.multiply(ca, cb, rlk, out=out); torch.cuda.synchronize()
    return g, ca, cb, rlk, out
def run(ctxs, streams, steps=6):
    for (g, ca, cb, rlk, out), s in zip(ctxs, streams):
        with torch.cuda.stream(s): g.multiply(ca, cb, rlk, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    for _ in range(steps):
        for (g, ca, cb, rlk, out), s in zip(ctxs, streams):
            with torch.cuda.stream(s): g.multiply(ca, cb, rlk, out=out)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
res = {}
total = int(sys.argv[1]) if len(sys.argv) > 1 else 8
one = mk(total, 100)
res["one_ctx"] = total / run([one], [torch.cuda.Stream()]) * 1e3
del one; torch.cuda.empty_cache()
for K in (2, 3, 4):
    if total % K: continue
    ctxs = [mk(total // K, 200 + 10 * k) for k in range(K)]
    res[f"{K}_ctx_{K}_streams"] = total / run(ctxs, [torch.cuda.Stream() for _ in range(K)]) * 1e3
    # staggered start: stream k begins after a delay so that phases of different contexts do not line up
    del ctxs; torch.cuda.empty_cache()
print(json.dumps({"total_batch": total, **{k: round(v, 1) for k, v in res.items()}}))
