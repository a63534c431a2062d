"""Secondary metric (BASELINE.json): BFV HMult+relinearize ops/s at config 4 (N=2^16, L=24, aux 25, dnum=3, K=8).
Imported by bench.py (--with-hmult) or run directly:  python bench_hmult.py [--batch B] [--steps K]"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


import contextlib


@contextlib.contextmanager
def gpu_numa_affinity(gpu_index: int):
    """For the duration of the block the process runs on the CPUs NVML reports as local to the GPU, so that pinned host buffers
    allocated (and first touched) inside it live on the GPU's own NUMA node; the previous affinity is restored afterwards (the CPU
    baseline wants every core).  On a two-socket box a buffer on the far socket costs the end-to-end path ~40% (103 k instead of
    170 k limb-transforms/s measured).  Best effort: any failure leaves the affinity alone."""
    saved = None
    try:
        if os.environ.get("FHE_B200_BENCH_NUMA", "1") == "0":      # A/B switch
            raise RuntimeError("disabled")
        saved = os.sched_getaffinity(0)
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:        # CUDA and NVML may enumerate differently: go through the PCI address
            pr = torch.cuda.get_device_properties(gpu_index)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
    except Exception:
        pass
    try:
        yield
    finally:
        if saved is not None:
            try:
                os.sched_setaffinity(0, saved)
            except Exception:
                pass


def _floor(n, L, R, K, dnum, ms_per_op, sms=148, mhz=1965.0):
    logn = n.bit_length() - 1
    ntts = 4 * (L + R) + 3 * (L + R) + dnum * (L + K) + 2 * (L + K)
    alpha = L // dnum
    macs = 4 * L * R + 3 * L * R + 3 * R * L + dnum * alpha * (L + K - alpha) + 2 * K * L          # per coefficient
    lanes_per_s = sms * 4 * 32 * mhz * 1e6
    ntt_ms = ntts * (n // 2) * logn * 28.0 / lanes_per_s * 1e3
    conv_ms = macs * n * 16.0 / lanes_per_s * 1e3
    return {"limb_ntts": ntts, "ntt_ms": ntt_ms, "conversion_macs_per_coefficient": macs, "conversion_ms_if_imad": conv_ms,
            "floor_ms": ntt_ms + conv_ms, "frac": (ntt_ms + conv_ms) / ms_per_op}


def run_hmult(args, local_rank=0, preset="c4", batch=None, steps=None, dist=None, lite=False):
    """dist: an initialised torch.distributed module for the batch-sharded multi-GPU run (every rank multiplies its own `batch`
    ciphertext pairs with replicated keys -- no data-path collective; the timed region is bracketed by barriers, the time is the
    max over ranks and `value` the whole-job aggregate)."""
    import numpy as np
    import torch
    import fhe_b200
    from fhe_b200.engine import pinned_empty, to_device, to_host
    from fhe_b200.params import bfv_preset

    torch.cuda.set_device(local_rank)
    if dist is not None:
        try:
            from bench import bind_to_gpu_numa_node
            bind_to_gpu_numa_node(local_rank)
        except Exception:
            pass
    p = bfv_preset(preset)
    n, L, t = p["n"], p["L"], p["t"]
    B = batch or 4
    K = steps or max(3, min(getattr(args, "steps", 10), 10))
    g = fhe_b200.BfvContext(n, L, p["R"], p["K"], p["dnum"], t, p["primes"], p["sigma"], p["hamming_weight"], device=local_rank)
    sk, pk = g.keygen(1, 2)
    rlk = g.relinkey_gen(3, sk)
    rng = np.random.default_rng(5)
    m1 = rng.integers(0, t, (B, n), dtype=np.uint64); m2 = rng.integers(0, t, (B, n), dtype=np.uint64)
    ca = g.encrypt(10, to_device(m1), pk); cb = g.encrypt(100, to_device(m2), pk)
    out = torch.empty_like(ca)
    lib = fhe_b200.load_library()
    for _ in range(3):
        g.multiply(ca, cb, rlk, out=out)
    torch.cuda.synchronize()
    world = dist.get_world_size() if dist is not None else 1
    if dist is not None:
        dist.barrier(); torch.cuda.synchronize()
    l0 = lib.fhe_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        g.multiply(ca, cb, rlk, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax[0])
    launches = lib.fhe_b200_launch_count() - l0
    # per-kernel-class breakdown of one more pass (CUDA events around every launch)
    import ctypes as C
    lib.fhe_b200_profile_enable(1)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    g.multiply(ca, cb, rlk, out=out)
    e3.record()
    torch.cuda.synchronize()
    breakdown = {}
    for kind, name in enumerate(("ntt_tile_fwd", "ntt_tile_inv", "ntt_row_fwd", "ntt_row_inv", "lincomb", "tensor", "ks_inner")):
        n_l, t_ms, u = C.c_uint64(), C.c_double(), C.c_uint64()
        lib.fhe_b200_profile_read(kind, C.byref(n_l), C.byref(t_ms), C.byref(u))
        breakdown[name] = {"launches": n_l.value, "ms": round(t_ms.value, 4)}
    breakdown["whole_call_ms"] = round(e2.elapsed_time(e3), 4)
    lib.fhe_b200_profile_enable(0)
    if lite:            # multi-GPU lines: the multiply itself plus its correctness check, nothing else
        import oracle
        from fhe_b200.engine import to_host as _th
        dec = _th(g.decrypt(out, sk))
        q0 = p["primes"][0]
        r = oracle.negacyclic_mul_ntt(m1[0], m2[0], q0)
        signed = np.where(r > np.uint64(q0 // 2), r.astype(np.int64) - np.int64(q0), r.astype(np.int64))
        ok = bool(np.array_equal(dec[0], np.mod(signed, np.int64(t)).astype(np.uint64)))
        return {"metric": "BFV HMult+relinearize ops/s", "value": world * B * K / (ms / 1e3), "unit": "ops/s", "n_gpus": world,
                "batch_per_gpu": B, "steps": K, "scaling": "weak", "parallelism": f"batch-sharded x{world} (independent ciphertexts, no data-path collective)",
                "ms_per_op": ms / (B * K), "config": {"workload": f"config4: N={n}, L={L}, R={p['R']}, dnum={p['dnum']}, K={p['K']}, t={t}"},
                "decrypts_to_product": ok, "gpu_launches": int(launches), "kernel_ms_per_call": breakdown}
    # the other two operations of the path: public-key encryption and decryption of the same batch (device-resident, CUDA events)
    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    d_m1 = to_device(m1)
    enc_ms = timed(lambda: g.encrypt(10, d_m1, pk))
    dec_ms = timed(lambda: g.decrypt(out, sk))
    # the same call with one operand twice (squaring path: two polynomials extended and transformed instead of four), and the
    # two halves on their own (three-component product; relinearisation of it)
    sq_ms = timed(lambda: g.multiply(ca, ca, rlk))
    ct3 = g.multiply_no_relin(ca, cb)
    mul3_ms = timed(lambda: g.multiply_no_relin(ca, cb, out=ct3))
    out2 = torch.empty_like(ca)
    relin_ms = timed(lambda: g.relinearize(ct3, rlk, out=out2))
    gk = g.galoiskey_gen(31, 3, sk)                      # rotate_rows by one step: automorphism + the same hybrid key switch
    rot_ms = timed(lambda: g.apply_galois(ca, 3, gk))
    del gk
    # invariant noise budget (bits) of one fresh ciphertext, of a product and of a depth-2 product: evidence that the parameter
    # set leaves room (host-side diagnostic, not timed)
    depth2 = g.multiply(out[0:1].contiguous(), ca[0:1].contiguous(), rlk)
    nb = g.noise_budget(torch.cat([ca[0:1], out[0:1], depth2]).contiguous(), sk)
    # correctness of what was timed: decrypt one result
    import oracle
    dec = to_host(g.decrypt(out, sk))
    q0 = p["primes"][0]                      # |coeff of the integer product| < N t^2 < q0/2: centred lift is exact
    r = oracle.negacyclic_mul_ntt(m1[0], m2[0], q0)
    signed = np.where(r > np.uint64(q0 // 2), r.astype(np.int64) - np.int64(q0), r.astype(np.int64))
    ok = bool(np.array_equal(dec[0], np.mod(signed, np.int64(t)).astype(np.uint64)))
    # end to end with host buffers (allocated on the GPU's NUMA node)
    with gpu_numa_affinity(local_rank):
        ha, hb = pinned_empty(tuple(ca.shape)), pinned_empty(tuple(cb.shape))
        ha[...] = to_host(ca); hb[...] = to_host(cb)
        ho = pinned_empty(tuple(ca.shape))
        ho[...] = 0
        g.multiply_host(ha, hb, rlk, ho)
        t0 = time.perf_counter()
        e2e_steps = 3
        for _ in range(e2e_steps):
            g.multiply_host(ha, hb, rlk, ho)
        e2e = time.perf_counter() - t0
    ok2 = bool(np.array_equal(ho, to_host(out)))
    ct_bytes = 2 * L * n * 8
    return {"metric": "BFV HMult+relinearize ops/s", "value": world * B * K / (ms / 1e3), "unit": "ops/s", "n_gpus": world,
            "batch_per_gpu": B, "batch": B, "steps": K, "scaling": "weak", "parallelism": f"batch-sharded x{world}",
            "ms_per_op": ms / (B * K), "config": {"workload": f"config4: N={n}, L={L}, R={p['R']}, dnum={p['dnum']}, K={p['K']}, t={t}"},
            "decrypts_to_product": ok, "gpu_launches": int(launches), "kernel_ms_per_call": breakdown,
            "e2e": {"value": B * e2e_steps / e2e, "unit": "ops/s", "h2d_bytes_per_step": 2 * B * ct_bytes, "d2h_bytes_per_step": B * ct_bytes,
                    "matches_device_path": ok2},
            "limb_ntts_per_op": 4 * (L + p["R"]) + 3 * (L + p["R"]) + p["dnum"] * (L + p["K"]) + 2 * (L + p["K"]),
            # integer-pipe floor of one multiply: limb-NTTs at 28 FMA-pipe cycles per warp-butterfly plus the
            # conversions at four IMAD.WIDE per (source, target) pair -- the second term is what the tensor-core kernel removes
            "int_pipe_floor": _floor(n, L, p["R"], p["K"], p["dnum"], ms / (B * K)),
            "encrypt": {"value": B / (enc_ms / 1e3), "unit": "ops/s", "ms_per_op": enc_ms / B, "limb_ntts_per_op": 3 * L},
            "decrypt": {"value": B / (dec_ms / 1e3), "unit": "ops/s", "ms_per_op": dec_ms / B, "limb_ntts_per_op": 2 * L},
            "noise_budget_bits": {"fresh": round(float(nb[0]), 2), "after_multiply": round(float(nb[1]), 2), "after_depth_2": round(float(nb[2]), 2)},
            "square": {"value": B / (sq_ms / 1e3), "unit": "ops/s", "ms_per_op": sq_ms / B},
            "multiply_no_relin": {"value": B / (mul3_ms / 1e3), "unit": "ops/s", "ms_per_op": mul3_ms / B},
            "relinearize": {"value": B / (relin_ms / 1e3), "unit": "ops/s", "ms_per_op": relin_ms / B},
            "rotate": {"value": B / (rot_ms / 1e3), "unit": "ops/s", "ms_per_op": rot_ms / B}}


def run_hmult_limb_sharded(local_rank=0, preset="c4", batch=1, steps=20, dist=None, ctx=None, keys=None):
    """ONE batch of ciphertext pairs multiplied by ALL ranks together (BASELINE.json config 4: limbs sharded over the GPUs):
    fhe_b200_bfv_multiply_relin_sharded, NTT-domain work limb-sharded, base conversions coefficient-sharded, the four
    transpositions per multiply stored by the producing kernels straight into the peers' buffers over NVLink (csrc/shard.cu).
    Timed with CUDA events on every rank between barriers, max over ranks; every rank checks its coefficient block against the
    plain single-GPU multiply of the same inputs, word for word."""
    import numpy as np
    import torch
    import fhe_b200
    from fhe_b200.engine import to_device
    from fhe_b200.params import bfv_preset

    torch.cuda.set_device(local_rank)
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    p = bfv_preset(preset)
    n, L, t = p["n"], p["L"], p["t"]
    g = ctx or fhe_b200.BfvContext(n, L, p["R"], p["K"], p["dnum"], t, p["primes"], p["sigma"], p["hamming_weight"], device=local_rank)
    if keys is None:
        sk, pk = g.keygen(1, 2)
        rlk = g.relinkey_gen(3, sk)
    else:
        sk, pk, rlk = keys
    rng = np.random.default_rng(5)
    B = batch
    m1 = rng.integers(0, t, (B, n), dtype=np.uint64); m2 = rng.integers(0, t, (B, n), dtype=np.uint64)
    dev = f"cuda:{local_rank}"
    ca = g.encrypt(10, to_device(m1, dev), pk); cb = g.encrypt(100, to_device(m2, dev), pk)
    want = g.multiply(ca, cb, rlk)
    # single-GPU rate of the same batch on this box (the denominator of the speed-up)
    for _ in range(2):
        g.multiply(ca, cb, rlk, out=want)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.multiply(ca, cb, rlk, out=want)
    e1.record(); torch.cuda.synchronize()
    single_ms = e0.elapsed_time(e1) / steps
    sh = fhe_b200.BfvShard(g, rank, world, max_batch=B)
    sh.connect() if dist is not None else sh.connect([sh.handle()])
    ks = sh.slice_key(rlk)
    sa, sb = sh.shard_ct(ca), sh.shard_ct(cb)
    out = torch.empty_like(sa)
    lib = fhe_b200.load_library()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()                 # a real stream: the library replays a multiply as a CUDA graph from the second identical call on
    with torch.cuda.stream(side):
        for _ in range(3):
            sh.multiply(sa, sb, ks, out=out)
        sh.check()
        if dist is not None:
            dist.barrier(); torch.cuda.synchronize()
        l0 = lib.fhe_b200_launch_count()
        e0.record()
        for _ in range(steps):
            sh.multiply(sa, sb, ks, out=out)
        e1.record()
        sh.check()
    ms = e0.elapsed_time(e1)
    launches = lib.fhe_b200_launch_count() - l0
    ok = bool(torch.equal(out, sh.shard_ct(want)))
    stat = torch.tensor([ms, 0.0 if ok else 1.0, float(sh.nvlink_bytes_per_op)], dtype=torch.float64, device=dev)
    tot = stat.clone()
    if dist is not None:
        dist.all_reduce(stat, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(stat[0])
    per_op_s = ms / 1e3 / steps / B
    nv_rank = float(stat[2]) * 1.0                     # bytes the busiest rank stores into its peers per ciphertext pair
    res = {"metric": "BFV HMult+relinearize ops/s", "value": B * steps / (ms / 1e3), "unit": "ops/s", "n_gpus": world, "batch": B,
           "steps": steps, "scaling": "strong", "parallelism": f"limb-sharded x{world} (one batch of {B} ciphertext pair(s) on all GPUs)",
           "ms_per_op": ms / (B * steps), "single_gpu_same_batch_ops_s": B / (single_ms / 1e3),
           "speedup_vs_single_gpu_same_batch": (single_ms / B) / (ms / (B * steps)),
           "matches_single_gpu_bit_exact": float(stat[1]) == 0.0, "launches_per_op_per_rank": launches / steps,
           "launch_note": "1 = one CUDA graph of the 27 kernels of a rank's multiply (FHE_B200_SHARD_GRAPH=0: kernel by kernel)",
           "config": {"workload": f"config4: N={n}, L={L}, R={p['R']}, dnum={p['dnum']}, K={p['K']}, t={t}"},
           "nvlink": {"bytes_per_op_all_ranks": float(tot[2]), "bytes_per_op_busiest_rank": nv_rank,
                      "achieved_GBs_per_rank_out": nv_rank / per_op_s / 1e9, "peak_GBs_per_direction": 900.0,
                      "frac": nv_rank / per_op_s / 1e9 / 900.0,
                      "note": "stores of the producing kernels into peer buffers (no separate transfer step, no NCCL on the data path); "
                              "an all-gather of the source limbs would move world x these bytes"}}
    sh.close()
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--preset", default="c4")
    ap.add_argument("--limb-sharded", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:        # torchrun: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench_hmult.py
        import torch
        import torch.distributed as dist
        lr = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(lr)
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        res = (run_hmult_limb_sharded(lr, a.preset, a.batch, a.steps, dist=dist) if a.limb_sharded
               else run_hmult(a, lr, a.preset, a.batch, a.steps, dist=dist))
        if dist.get_rank() == 0:
            print(json.dumps(res), flush=True)
        dist.barrier()
        dist.destroy_process_group()
    else:
        print(json.dumps(run_hmult_limb_sharded(0, a.preset, a.batch, a.steps) if a.limb_sharded else run_hmult(a, 0, a.preset, a.batch, a.steps)))
