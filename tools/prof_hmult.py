"""One BFV multiply+relinearize at config 4 for ncu launch lists: python tools/prof_hmult.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fhe_b200
from fhe_b200.engine import to_device
from fhe_b200.params import bfv_preset
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
p = bfv_preset("c4"); n, t = p["n"], p["t"]
g = fhe_b200.BfvContext(n, p["L"], p["R"], p["K"], p["dnum"], t, p["primes"], p["sigma"], p["hamming_weight"])
sk, pk = g.keygen(1, 2); rlk = g.relinkey_gen(3, sk)
rng = np.random.default_rng(5)
ca = g.encrypt(10, to_device(rng.integers(0, t, (B, n), dtype=np.uint64)), pk)
cb = g.encrypt(100, to_device(rng.integers(0, t, (B, n), dtype=np.uint64)), pk)
out = torch.empty_like(ca)
g.multiply(ca, cb, rlk, out=out); torch.cuda.synchronize()
torch.cuda.nvtx.range_push("hmult")
g.multiply(ca, cb, rlk, out=out); torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("prof_hmult ok")
