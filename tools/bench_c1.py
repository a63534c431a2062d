"""BASELINE.json config 1 latency: N = 4096, one 60-bit prime, one polynomial: c = INTT(NTT(a) . NTT(b)), fused kernel vs four launches.
    python tools/bench_c1.py"""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch, fhe_b200
    from fhe_b200.engine import to_device
    from fhe_b200.params import prime_chain
    q = prime_chain(1)[0]; n = 4096
    plan = fhe_b200.Plan(n, [q])
    rng = np.random.default_rng(1)
    a = to_device(rng.integers(0, q, (1, 1, n), dtype=np.uint64)); b = to_device(rng.integers(0, q, (1, 1, n), dtype=np.uint64))
    out = torch.empty_like(a)
    for _ in range(20): plan.negacyclic_mul(a, b, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2000
    e0.record()
    for _ in range(reps): plan.negacyclic_mul(a, b, out=out)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"config": "C1: N=4096, 1 limb, 1 polynomial", "fused": os.environ.get("FHE_B200_MUL_FUSED"), "us_per_op": e0.elapsed_time(e1) * 1e3 / reps}))
else:
    for f in ("1", "0"):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, FHE_B200_MUL_FUSED=f))
