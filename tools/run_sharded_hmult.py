"""Limb-sharded BFV multiply+relinearize across the GPUs of one node (BASELINE.json config 4).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29544 \
        tools/run_sharded_hmult.py [--preset c4] [--steps 10]
Every rank builds the same keys and ciphertexts (deterministic seeds), keeps only its limbs, runs the sharded
multiply (NCCL all-gathers before each base conversion), and rank 0 checks the gathered result against the
single-GPU fhe_b200_bfv_multiply_relin bit for bit.  Prints one JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import fhe_b200
import fhe_b200.parallel as par
from fhe_b200.engine import to_device
from fhe_b200.params import bfv_preset

ap = argparse.ArgumentParser()
ap.add_argument("--preset", default="c4")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--profile", action="store_true", help="rank 0 prints the kernels of three multiplies by CUDA time (torch.profiler)")
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
p = bfv_preset(a.preset)
n, L, R, K, dnum, t = p["n"], p["L"], p["R"], p["K"], p["dnum"], p["t"]
g = fhe_b200.BfvContext(n, L, R, K, dnum, t, p["primes"], p["sigma"], p["hamming_weight"], device=lr)
sk, pk = g.keygen(1, 2); rlk = g.relinkey_gen(3, sk)
rng = np.random.default_rng(4)
ca = g.encrypt(5, to_device(rng.integers(0, t, (1, n), dtype=np.uint64), torch.device("cuda", lr)), pk)
cb = g.encrypt(6, to_device(rng.integers(0, t, (1, n), dtype=np.uint64), torch.device("cuda", lr)), pk)
expect = g.multiply(ca, cb, rlk)[0]
be = par.GpuBackend(n, p["primes"], lr)
sb = par.LimbShardedBfv(n, L, R, K, dnum, t, p["primes"], be, rank=rank, world=world)
kq, kp = sb.shard_relin_key(rlk)
al, bl = sb.shard_ciphertext(ca[0]), sb.shard_ciphertext(cb[0])
out = sb.multiply_relin(al, bl, kq, kp)
ok = bool(torch.equal(sb.gather_ciphertext(out), expect))
for _ in range(2):
    sb.multiply_relin(al, bl, kq, kp)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    sb.multiply_relin(al, bl, kq, kp)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=torch.device("cuda", lr))
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if a.profile:
    import time
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        sb.multiply_relin(al, bl, kq, kp)
    t_issue = (time.perf_counter() - t0) / 3
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            sb.multiply_relin(al, bl, kq, kp)
        torch.cuda.synchronize()
    if rank == 0:
        print(f"host time to issue one multiply: {t_issue * 1e3:.3f} ms", flush=True)
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60), flush=True)
# single-GPU time of the same op on rank 0 for the scaling ratio
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
for _ in range(a.steps):
    g.multiply(ca, cb, rlk)
s1.record(); torch.cuda.synchronize()
if rank == 0:
    print(json.dumps({"metric": "BFV HMult+relinearize ops/s (limb-sharded, one ciphertext)", "n_gpus": world, "preset": a.preset,
                      "value": a.steps / (float(ms[0]) / 1e3), "ms_per_op": float(ms[0]) / a.steps,
                      "single_gpu_fused_ms_per_op": s0.elapsed_time(s1) / a.steps,
                      "bit_exact_vs_single_gpu": ok, "allgather_bytes_received_per_rank_per_op": sb.gather_bytes_per_op(),
                      "collectives_per_op": 5}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
