#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 RNS-NTT engine (contract: see the task's "Measurement" section).

Metric (BASELINE.json): batched RNS-NTT limb-transforms/s at N=2^16.  Workload = BASELINE.json config 3 per GPU:
uint64 [64 polynomials][32 limbs of ~60-bit primes][65536], one step = forward NTT of all + inverse NTT of all
= 4096 limb-transforms (one limb-transform = one N-point forward OR inverse negacyclic NTT of one limb of one
polynomial; algorithmic HBM bytes 2*N*8 = 1 MiB).  N GPUs: every rank owns its own 64 polynomials
(batch sharding, no data-path collective) -> weak scaling.  The secondary metric (BFV HMult+relinearize ops/s,
BASELINE.json config 4) is reported under "hmult" when --with-hmult (default on at N=1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOGN = 16
N = 1 << LOGN
LIMBS = 32
POLYS = 64
UNIT_BYTES = 2 * N * 8                       # algorithmic bytes per limb-transform (read once, write once)
UNITS_PER_STEP = 2 * LIMBS * POLYS           # forward + inverse over the whole batch
METRIC = "batched RNS-NTT limb-transforms/s at N=2^16"
UNIT = "limb-transforms/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """Times the CPU restatement (oracle port; the reference's own code for this path neither builds nor computes
    a transform) on the box's host cores, all threads, bounded sample per step."""
    if rank != 0:
        return
    import numpy as np
    import oracle
    oracle.build()
    chain = oracle.prime_chain(LIMBS)
    eng = oracle.RnsNtt(N, chain)
    # torchrun exports OMP_NUM_THREADS=1; the reference arm is meant to use every host core it can
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sample_polys = max(1, min(POLYS, threads // 8 if threads >= 16 else 1))
    rng = np.random.default_rng(0x5EED0003)
    data = np.stack([np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in chain]) for _ in range(sample_polys)])
    flat = data.reshape(-1)
    units = 2 * LIMBS * sample_polys

    def step():
        eng.run_inplace(flat, sample_polys, False, threads)
        eng.run_inplace(flat, sample_polys, True, threads)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = units * args.steps / dt
    sample = f"{sample_polys} of {POLYS} polynomials x {LIMBS} limbs, forward+inverse ({units} limb-transforms per step)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "config3: uint64[64][32][65536] fwd+inv RNS-NTT (bounded sample per step)",
                   "N": N, "limbs": LIMBS, "polys": POLYS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference CUDA code does not build/compute (SURVEY section 0); this arm is the CPU oracle port, OpenMP over polynomial x limb",
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample():
    """bounded (~10 s) CPU sample of the same workload for the ours-arm JSON line (rank 0, N=1 only)."""
    import numpy as np
    import oracle
    oracle.build()
    chain = oracle.prime_chain(LIMBS)
    eng = oracle.RnsNtt(N, chain)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    rng = np.random.default_rng(0x5EED0003)
    polys = 1
    data = np.stack([np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in chain]) for _ in range(polys)])
    flat = data.reshape(-1)
    eng.run_inplace(flat, polys, False, threads); eng.run_inplace(flat, polys, True, threads)          # warm
    t0 = time.perf_counter(); reps = 0
    while True:
        eng.run_inplace(flat, polys, False, threads); eng.run_inplace(flat, polys, True, threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt > 8.0 or reps >= 200:
            break
    units = reps * polys * 2 * LIMBS
    return {"value": units / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{reps} x (1 polynomial x {LIMBS} limbs, forward+inverse) = {units} limb-transforms in {dt:.1f} s, "
                      f"oracle Shoup/Harvey NTT with OpenMP"}


def bind_to_gpu_numa_node(gpu_index: int):
    """pins this process to the CPUs NVML reports as local to the GPU, so that the pinned host buffers of the end-to-end path
    are allocated on (and copied from) the GPU's own NUMA node.  Without it the eight ranks of a node share one socket's
    memory and root complex (measured: 31 k limb-transforms/s per GPU end to end instead of 174 k).  Best effort."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:        # CUDA and NVML may enumerate differently: go through the PCI address
            pr = torch.cuda.get_device_properties(gpu_index)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False


# --------------------------------------------------------------------------------------------- ours
def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import fhe_b200
    from fhe_b200.engine import pinned_empty

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:            # (single-GPU runs keep every host core: the CPU baseline of the same line uses them)
        bind_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"        # keep stdout to the one JSON line (NCCL prints its version banner there)
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    lib = fhe_b200.load_library()

    # deterministic prime chain, computed by the library's own host code via a throw-away import-free path
    from fhe_b200.params import prime_chain
    chain = prime_chain(LIMBS)
    plan = fhe_b200.Plan(N, chain, device=local_rank)

    g = torch.Generator(device=dev); g.manual_seed(0x5EED0003 + rank)
    x = torch.empty((POLYS, LIMBS, N), dtype=torch.int64, device=dev)
    for l, q in enumerate(chain):
        x[:, l, :] = torch.randint(0, q, (POLYS, N), generator=g, device=dev, dtype=torch.int64)
    ref = x.clone()

    def step():
        plan.forward(x)
        plan.inverse(x)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.fhe_b200_launch_count()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    launches = lib.fhe_b200_launch_count() - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    ok = bool(torch.equal(x, ref))                       # forward+inverse is the identity: free correctness check

    # per-kernel durations for the roofline (second pass, events around every launch, same stream)
    lib.fhe_b200_profile_enable(1)
    psteps = min(args.steps, 5)
    for _ in range(psteps):
        step()
    torch.cuda.synchronize()
    kinds = {}
    for kind, name in enumerate(("tile_fwd", "tile_inv", "row_fwd", "row_inv")):
        n_l, t_ms, u = C.c_uint64(), C.c_double(), C.c_uint64()
        lib.fhe_b200_profile_read(kind, C.byref(n_l), C.byref(t_ms), C.byref(u))
        kinds[name] = (n_l.value, t_ms.value, u.value)
    lib.fhe_b200_profile_enable(0)

    # end to end through the host-buffer entry point (pinned host memory, H2D + D2H inside the timed region)
    from bench_hmult import gpu_numa_affinity
    with gpu_numa_affinity(local_rank):                      # the pinned buffer is allocated and first touched on the GPU's NUMA node
        h = pinned_empty((POLYS, LIMBS, N))
        h[...] = ref.cpu().numpy().view(np.uint64)
        e2e_steps = max(2, min(args.steps, 4))
        plan.ntt_host(h, 2)                                  # warm-up (allocates staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.ntt_host(h, 2)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    e2e_ok = bool(np.array_equal(h, ref.cpu().numpy().view(np.uint64)))

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])

    if rank == 0:
        peak, peak_src = measured_peaks()
        value = world * UNITS_PER_STEP * args.steps / (ms_max / 1e3)
        e2e_val = world * UNITS_PER_STEP * e2e_steps / (e2e_ms_max / 1e3)
        # N = 2^16 runs as two balanced passes per direction (8 stages each; ntt_bal.cu): every launch reads and writes each limb
        # once, i.e. moves the algorithmic 1 MiB per limb-transform it covers.  The dominant kernel is the slowest of the four.
        names = {"tile_fwd": "bal_b_kernel<8,16,near,fwd> (tile-pair pass)", "tile_inv": "bal_b_kernel<8,16,near,inv> (tile-pair pass)",
                 "row_fwd": "bal_a_kernel<8,16,near,fwd> (column pass)", "row_inv": "bal_a_kernel<8,16,near,inv> (column pass)"}
        if os.environ.get("FHE_B200_NTT_BAL", "1") == "0":
            names = {"tile_fwd": "ntt_tile_fwd_kernel<12,4,16>", "tile_inv": "ntt_tile_inv_kernel<12,4,16>",
                     "row_fwd": "ntt_row_fwd_kernel<12,4,16>", "row_inv": "ntt_row_inv_kernel<12,4,16>"}
        dom = max(kinds, key=lambda k: kinds[k][1])
        tl, tms, tu = kinds[dom]
        all_ms = sum(v[1] for v in kinds.values())
        achieved = (tu * UNIT_BYTES / 1e9) / (tms / 1e3) if tms > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch", {}).get(dom)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": "config3: uint64[64][32][65536] per GPU, forward then inverse batched RNS-NTT "
                                   "(4096 limb-transforms per step per GPU)",
                       "N": N, "limbs": LIMBS, "polys_per_gpu": POLYS, "parallelism": f"batch-sharded x{world}",
                       "l2_policy": "inputs (1 GiB per GPU) exceed the 126 MB L2; no flush needed",
                       "ntt_chunk_mb": int(os.environ.get("FHE_B200_NTT_CHUNK_MB", "1024"))},
            "roundtrip_bit_exact": ok and e2e_ok,
            "clocks": clocks,
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": POLYS * LIMBS * N * 8,
                    "d2h_bytes_per_step": POLYS * LIMBS * N * 8, "steps": e2e_steps,
                    "api": "fhe_b200_ntt_host (pinned host buffer, 3-stream chunk pipeline)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "kernel": names[dom], "peak_source": peak_src,
                         "launches_profiled": tl, "avg_launch_ms": (tms / tl) if tl else None,
                         "units_per_launch": (tu / tl) if tl else None, "bytes_per_unit": UNIT_BYTES,
                         "kernel_share_of_step": (tms / all_ms) if all_ms else None,
                         "step_achieved_GBs": value * UNIT_BYTES / 1e9 / world,
                         "step_frac": value * UNIT_BYTES / 1e9 / world / peak,
                         "per_kernel_ms": {k: v[1] / psteps for k, v in kinds.items()}},
        }
        # integer-pipe roofline (the binding one): 28 FMA-pipe cycles per warp-butterfly
        # (5 IMAD.WIDE at 4 cycles, 4 IMAD.lo at 2 cycles per warp instruction per SM sub-partition; measured by
        # tools/microbench2.cu), 4 sub-partitions per SM, at the SM clock sampled during the timed region
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        peak_bfly = 148 * 4 * 32 * sm_mhz * 1e6 / 28.0
        bfly = value / world * (N // 2) * LOGN
        line["int_pipe"] = {"bound": "imad", "achieved_Gbutterflies_s": bfly / 1e9, "peak_Gbutterflies_s": peak_bfly / 1e9,
                            "frac": bfly / peak_bfly, "pipe_cycles_per_warp_butterfly": 28, "sm_mhz": sm_mhz,
                            "peak_limb_transforms_s": peak_bfly / ((N // 2) * LOGN)}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample()
        if world == 1 and not args.no_hmult:
            try:
                from bench_hmult import run_hmult
                line["hmult"] = run_hmult(args, local_rank, batch=8)
            except Exception as ex:  # secondary metric must never take the headline line down
                line["hmult"] = {"error": repr(ex)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--with-hmult", action="store_true", default=False)
    ap.add_argument("--no-hmult", action="store_true", default=False)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
