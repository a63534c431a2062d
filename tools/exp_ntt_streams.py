"""Experiment: config 3 (64 polynomials x 32 limbs, N = 2^16) forward + inverse as one call vs K slices of the batch on K streams."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fhe_b200, oracle
n, limbs, polys = 1 << 16, 32, 64
plan = fhe_b200.Plan(n, oracle.prime_chain(limbs))
x = torch.randint(0, 1 << 59, (polys, limbs, n), dtype=torch.int64, device="cuda")
def run(K, steps=10):
    streams = [torch.cuda.Stream() for _ in range(K)]
    parts = list(x.chunk(K, dim=0))
    def once():
        for s in streams: s.wait_stream(torch.cuda.current_stream())
        for p, s in zip(parts, streams):
            with torch.cuda.stream(s):
                plan.forward(p); plan.inverse(p)
        for s in streams: torch.cuda.current_stream().wait_stream(s)
    for _ in range(3): once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): once()
    e1.record(); torch.cuda.synchronize()
    return 2 * polys * limbs * steps / (e0.elapsed_time(e1) / 1e3)
print(json.dumps({f"{K}_streams": round(run(K)) for K in (1, 2, 4)}))
