"""Base conversions of one config-4 multiply on their own: CUDA-event time per shape for the tcgen05 and the mma.sync kernels, or
(PROF=1) a short run for ncu.   python tools/prof_lincomb.py [batch]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
prof = os.environ.get("PROF") == "1"
n = 1 << 16
res = {}
for kern in (("tc",) if prof else ("tc", "mma")):
    os.environ["FHE_B200_LINCOMB_TC"] = "1" if kern == "tc" else "0"
    import fhe_b200
    from fhe_b200.params import prime_chain
    chain = prime_chain(49)
    Q, R = chain[:24], chain[24:]
    shapes = {"q2r 24->25 (x4 per multiply)": (fhe_b200.LinComb.conv(Q, R), Q, B, None),
              "scale 24(+25)->25 (x3)": (fhe_b200.LinComb.scale(Q, R, 65537, R, True), Q, 3 * B, R),
              "r2q 25->24 (x3)": (fhe_b200.LinComb.conv(R, Q), R, 3 * B, None),
              "modup 8->24 (x3)": (fhe_b200.LinComb.conv(Q[:8], Q[8:] + R[:8]), Q[:8], B, None)}
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for name, (lc, src, polys, extra) in shapes.items():
        x = torch.stack([torch.randint(0, q, (polys, n), generator=g, device="cuda", dtype=torch.int64) for q in src], dim=1).contiguous()
        ex = None if extra is None else torch.stack([torch.randint(0, q, (polys, n), generator=g, device="cuda", dtype=torch.int64) for q in extra], dim=1).contiguous()
        y = lc.apply(x, ex)
        torch.cuda.synchronize()
        if prof:
            lc.apply(x, ex); torch.cuda.synchronize()
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 100
        for _ in range(10):
            lc.apply(x, ex)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            lc.apply(x, ex)
        e1.record(); torch.cuda.synchronize()
        res.setdefault(name, {})[kern] = round(e0.elapsed_time(e1) / reps * 1e3, 1)
        res[name].setdefault("polys", polys)
        res[name]["sha_" + kern] = int(y.sum().item()) & 0xffffffff
print(json.dumps(res, indent=1))
print("prof_lincomb ok")
