"""Small, deterministic invocation of the batched RNS-NTT for ncu captures:
   python tools/prof_ntt.py [polys] [reps]   -> forward+inverse over uint64[polys][32][65536], `reps` times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fhe_b200
from fhe_b200.params import prime_chain

polys = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
N, LIMBS = 1 << 16, 32
chain = prime_chain(LIMBS)
plan = fhe_b200.Plan(N, chain)
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.empty((polys, LIMBS, N), dtype=torch.int64, device="cuda")
for l, q in enumerate(chain):
    x[:, l, :] = torch.randint(0, q, (polys, N), generator=g, device="cuda", dtype=torch.int64)
ref = x.clone()
for _ in range(reps):
    plan.forward(x)
    plan.inverse(x)
torch.cuda.synchronize()
assert torch.equal(x, ref)
print("prof_ntt ok", polys, reps)
