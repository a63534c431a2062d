"""BASELINE.json config 5: throughput sweep N = 2^12..2^17 x batch 1..1024 ciphertexts at 1/2/4/8 B200, next to the reference's own
CUDA kernels and the host CPU baseline on the same box.

    python tools/sweep_c5.py [--quick] [--out gpurun_out/sweep_c5_1.jsonl]                 # one GPU
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/sweep_c5.py ...   # G GPUs, batch-sharded

Per point (N, L(N), batch): M1 = forward+inverse RNS-NTT of the batch's 2*batch polynomials (limb-transforms/s) and M2 = BFV
HMult+relinearize of `batch` ciphertext pairs (ops/s; parameters L(N), R = L+1, dnum/K as the compat FHEContext chooses them).  On
G GPUs every rank runs the same batch on its own ciphertexts (batch sharding: no data-path collective), the time is the max over
ranks and the rate the whole-job aggregate.  Columns on rank 0 of the one-GPU run only: the IMAD-mix roofline fraction of M1 (peaks
measured in this run, as bench.py does), the CPU port's M1 rate at that N (all host threads), and per-coefficient times of the
reference's ntt_pointwise_mul_kernel / poly_add_kernel (256-bit elements) against the engine's poly_mul / poly_add (64-bit), plus
the reference's NTTEngine::forward at N = 1024 -- the only size it can launch (src/ntt.cu:36-39)."""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import fhe_b200
from fhe_b200.params import prime_chain

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true", help="N in {2^13, 2^16}, batches {1, 16, 256} (multi-GPU runs)")
ap.add_argument("--out", default=None)
ap.add_argument("--max-bytes", type=float, default=24e9)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist = None
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
lib = fhe_b200.load_library()
LIMBS = {12: 2, 13: 4, 14: 8, 15: 16, 16: 24, 17: 32}
LOGNS = (13, 16) if a.quick else (12, 13, 14, 15, 16, 17)
BATCHES = (1, 16, 256) if a.quick else (1, 4, 16, 64, 256, 1024)
chain = prime_chain(65)
out = open(a.out, "w") if (a.out and rank == 0) else None


def emit(d):
    if rank == 0:
        s = json.dumps(d)
        print(s, flush=True)
        if out:
            out.write(s + "\n"); out.flush()


def max_over_ranks(ms):
    if dist is None:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    if dist is not None:
        dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return max_over_ranks(e0.elapsed_time(e1) / reps)


lo, wide = C.c_double(), C.c_double()
fhe_b200.check(lib.fhe_b200_measure_int_peaks(lr, C.byref(lo), C.byref(wide), 5))
mix_peak = 9.0 / (5.0 / wide.value + 4.0 / lo.value)          # IMAD-class ops/s of the butterfly's 5:4 mix (bench.py)
emit({"what": "peaks", "n_gpus": world, "imad_lo_Tops": lo.value / 1e12, "imad_wide_Tops": wide.value / 1e12, "mix_peak_Tops": mix_peak / 1e12})

one = world == 1
oracle = None
if one:
    import oracle as _o
    oracle = _o
    oracle.use_native() or oracle.build()
    threads = len(os.sched_getaffinity(0))

for logn in LOGNS:
    n, L = 1 << logn, LIMBS[logn]
    plan = fhe_b200.Plan(n, chain[:L], device=lr)
    # ---- host CPU baseline and reference kernels at this N (one GPU run only)
    cpu_rate = None
    if one:
        eng = oracle.RnsNtt(n, chain[:L])
        polys = max(1, min(64, (threads * 4) // L + 1))
        flat = np.concatenate([np.random.default_rng(logn).integers(0, q, n, dtype=np.uint64) for q in chain[:L]] * polys)
        eng.run_inplace(flat, polys, False, threads); eng.run_inplace(flat, polys, True, threads)
        t0 = time.perf_counter(); r = 0
        while time.perf_counter() - t0 < 1.0:
            eng.run_inplace(flat, polys, False, threads); eng.run_inplace(flat, polys, True, threads); r += 1
        cpu_rate = r * 2 * polys * L / (time.perf_counter() - t0)
        ref_mul = oracle.ref_time_poly("pointwise_mul", chain[0], n, 50)
        ref_add = oracle.ref_time_poly("poly_add", chain[0], n, 50)
        x1 = torch.randint(0, chain[0], (64, 1, n), device=dev, dtype=torch.int64); y1 = torch.empty_like(x1)
        p1 = fhe_b200.Plan(n, chain[:1], device=lr)
        ours_mul = timed(lambda: p1.mul(x1, x1, out=y1), 50) / 64          # per polynomial of n coefficients (64 in one launch)
        ours_add = timed(lambda: p1.add(x1, x1, out=y1), 50) / 64
        ours_mul1 = timed(lambda: p1.mul(x1[:1], x1[:1], out=y1[:1]), 50)  # one polynomial per launch, as the reference launches
        emit({"what": "reference_kernels", "logn": logn,
              "ref_ntt_pointwise_mul_us": None if ref_mul is None else ref_mul * 1e3, "ref_poly_add_us": None if ref_add is None else ref_add * 1e3,
              "ours_poly_mul_us_one_launch_per_poly": ours_mul1 * 1e3, "ours_poly_mul_us_per_poly_batched64": ours_mul * 1e3,
              "ours_poly_add_us_per_poly_batched64": ours_add * 1e3,
              "note": "reference: 256-bit elements, one polynomial of n coefficients per launch (src/ntt.cu:65-68, src/polynomial.cu:36-42)"})
        del p1, x1, y1
    for batch in BATCHES:
        polys = 2 * batch
        if polys * L * n * 8 > a.max_bytes:
            continue
        g = torch.Generator(device=dev); g.manual_seed(logn * 1000 + batch + rank)
        x = torch.empty((polys, L, n), dtype=torch.int64, device=dev)
        for l, q in enumerate(chain[:L]):
            x[:, l, :] = torch.randint(0, q, (polys, n), generator=g, device=dev, dtype=torch.int64)
        ref = x.clone()
        units = 2 * polys * L
        reps = max(3, min(100, int(1e5 / units)))
        ms = timed(lambda: (plan.forward(x), plan.inverse(x)), reps)
        ok = bool(torch.equal(x, ref))
        rate = world * units / (ms / 1e3)
        bfly = rate / world * (n // 2) * logn
        emit({"what": "M1", "n_gpus": world, "logn": logn, "limbs": L, "ciphertexts_per_gpu": batch, "ms_fwd_inv": round(ms, 4),
              "limb_transforms_per_s": round(rate), "per_gpu_butterflies_G_s": round(bfly / 1e9, 1), "imad_mix_roofline_frac": round(bfly * 9 / mix_peak, 3),
              "cpu_port_limb_transforms_per_s": None if cpu_rate is None else round(cpu_rate), "cpu_threads": threads if one else None,
              "speedup_vs_cpu": None if cpu_rate is None else round(rate / cpu_rate, 1), "roundtrip_bit_exact": ok})
        del x, ref
    del plan
    # ---- M2: HMult + relinearize
    if 2 * L + 1 > 62:
        L = 30                                     # the library holds at most 62 primes per context: N = 2^17 runs HMult with L = 30, R = 31
    R = L + 1
    dnum = 3 if (L >= 6 and L % 3 == 0) else (2 if (L >= 4 and L % 2 == 0) else L)
    K = L // dnum
    if logn == 16:
        R, K, dnum = 25, 8, 3                      # BASELINE.json config 4
    t = 65537
    ctx = fhe_b200.BfvContext(n, L, R, K, dnum, t, chain[:L + R], 3.2, 64, device=lr)
    sk, pk = ctx.keygen(1, 2); rlk = ctx.relinkey_gen(3, sk)
    A, W = L + R, L + K
    per_ct = (7 * A + 3 * R + 3 * L + (dnum + 2) * W) * n * 8 + 6 * L * n * 8
    for batch in BATCHES:
        if batch * per_ct > a.max_bytes:
            continue
        m = torch.randint(0, t, (batch, n), device=dev, dtype=torch.int64)
        ca = ctx.encrypt(10 + rank, m, pk); cb = ctx.encrypt(100 + rank, m, pk)
        o = torch.empty_like(ca)
        reps = max(2, min(20, int(64 / batch) + 1))
        ms = timed(lambda: ctx.multiply(ca, cb, rlk, out=o), reps)
        # correctness of what was timed: decrypt(product) == m * m for the first ciphertext (host NTT product through the oracle on one GPU)
        ok = None
        if one and logn <= 16:
            dec = ctx.decrypt(o[:1].contiguous(), sk).cpu().numpy().view(np.uint64)[0]
            mm = m[0].cpu().numpy().view(np.uint64)
            q0 = chain[0]
            r = oracle.negacyclic_mul_ntt(mm, mm, q0)
            signed = np.where(r > np.uint64(q0 // 2), r.astype(np.int64) - np.int64(q0), r.astype(np.int64))
            ok = bool(np.array_equal(dec, np.mod(signed, np.int64(t)).astype(np.uint64)))
        emit({"what": "M2", "n_gpus": world, "logn": logn, "L": L, "R": R, "K": K, "dnum": dnum, "ciphertext_pairs_per_gpu": batch,
              "ms_per_op": round(ms / batch, 4), "hmult_relin_ops_per_s": round(world * batch / (ms / 1e3), 1), "decrypts_to_product": ok,
              "reference": "not implemented (relinearize is components.resize(2), src/fhe.cu:226-235; NTT launches fail for N > 1024, src/ntt.cu:36-39)"})
        del ca, cb, o, m
    del ctx, sk, pk, rlk
    torch.cuda.empty_cache()

if one:
    t1024 = oracle.ref_time_poly("ntt_forward", 12289, 1024, 50)
    p = fhe_b200.Plan(1024, [12289], device=lr)
    x = torch.randint(0, 12289, (1, 1, 1024), device=dev, dtype=torch.int64)
    ours1 = timed(lambda: p.forward(x), 100)
    xb = torch.randint(0, 12289, (4096, 1, 1024), device=dev, dtype=torch.int64)
    oursb = timed(lambda: p.forward(xb), 20) / 4096
    emit({"what": "reference_ntt_forward_n1024", "ref_us_per_transform": None if t1024 is None else t1024 * 1e3, "ours_us_one_transform": ours1 * 1e3,
          "ours_us_per_transform_batched4096": oursb * 1e3,
          "note": "the reference's transform (bit_reverse_kernel + one block of 1024 threads, 256-bit Montgomery, placeholder twiddles: output "
                  "meaningless, only timed); N = 1024 is the largest size it can launch"})
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
