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
WORKLOAD = ("config3: uint64[64][32][65536] per GPU, forward then inverse batched RNS-NTT "
            "(4096 limb-transforms per step per GPU)")


def config_dict(world):
    """identical in both arms (the reference arm runs the same workload on the host cores)"""
    return {"workload": WORKLOAD, "N": N, "limbs": LIMBS, "polys_per_gpu": POLYS, "parallelism": f"batch-sharded x{world}",
            "l2_policy": "inputs (1 GiB per GPU) exceed the 126 MB L2; no flush needed",
            "ntt_chunk_mb": int(os.environ.get("FHE_B200_NTT_CHUNK_MB", "1024"))}


def _host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _cpu_engine():
    """the CPU port (oracle: Harvey lazy butterflies, Shoup twiddles, OpenMP over polynomial x limb), compiled -O3 -march=native on
    this machine when gcc is there (oracle.use_native), else the portable build that ships with the snapshot"""
    import oracle
    native = oracle.use_native()
    if not native:
        oracle.build()
    chain = oracle.prime_chain(LIMBS)
    return oracle, oracle.RnsNtt(N, chain), chain, ("-O3 -march=native" if native else "-O3 -march=x86-64-v2")


def run_reference(args, rank, world):
    """Times the CPU restatement (oracle port; the reference's own code for this path neither builds nor computes
    a transform) on the box's host cores, all threads.  A step is the same workload as the GPU arm's -- all 64
    polynomials forward and inverse -- unless the host is so small that a step would take more than ~2 s, in which case a step
    is the largest whole number of polynomials that fits (stated in cpu_baseline.sample)."""
    if rank != 0:
        return
    import numpy as np
    oracle, eng, chain, flags = _cpu_engine()
    threads = _host_threads()       # torchrun exports OMP_NUM_THREADS=1; the reference arm is meant to use every host core it can
    rng = np.random.default_rng(0x5EED0003)
    one = np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in chain])
    probe = np.tile(one.reshape(-1), max(1, min(POLYS, threads)))
    pp = probe.size // (LIMBS * N)
    eng.run_inplace(probe, pp, False, threads); eng.run_inplace(probe, pp, True, threads)
    t0 = time.perf_counter()
    eng.run_inplace(probe, pp, False, threads); eng.run_inplace(probe, pp, True, threads)
    per_poly = (time.perf_counter() - t0) / pp
    sample_polys = int(max(1, min(POLYS, 2.0 / max(per_poly, 1e-6))))
    if sample_polys >= POLYS // 2:
        sample_polys = POLYS
    flat = np.tile(one.reshape(-1), sample_polys)
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
    ok = bool(np.array_equal(flat, np.tile(one.reshape(-1), sample_polys)))
    value = units * args.steps / dt
    sample = (f"{sample_polys} of {POLYS} polynomials x {LIMBS} limbs per step, forward+inverse ({units} limb-transforms per step), "
              f"{threads} OpenMP threads, gcc {flags}; {1e9 / (value / threads * (N // 2) * LOGN):.2f} ns per butterfly per thread")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_dict(world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "roundtrip_bit_exact": ok,
        "note": "reference CUDA code does not build/compute (SURVEY section 0); this arm is the CPU oracle port, OpenMP over polynomial x limb",
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample():
    """bounded (~10 s) CPU sample of the same workload for the ours-arm JSON line (rank 0, N=1 only)."""
    import numpy as np
    oracle, eng, chain, flags = _cpu_engine()
    threads = _host_threads()
    rng = np.random.default_rng(0x5EED0003)
    polys = max(1, min(POLYS, threads // 4))
    data = np.stack([np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in chain]) for _ in range(polys)])
    flat = data.reshape(-1)
    eng.run_inplace(flat, polys, False, threads); eng.run_inplace(flat, polys, True, threads)          # warm
    t0 = time.perf_counter(); reps = 0
    while True:
        eng.run_inplace(flat, polys, False, threads); eng.run_inplace(flat, polys, True, threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt > 10.0 or reps >= 400:
            break
    units = reps * polys * 2 * LIMBS
    rate = units / dt
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{reps} x ({polys} polynomial(s) x {LIMBS} limbs, forward+inverse) = {units} limb-transforms in {dt:.1f} s, "
                      f"oracle Harvey/Shoup NTT with OpenMP, gcc {flags}; {1e9 / (rate / threads * (N // 2) * LOGN):.2f} ns per butterfly per thread"}


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
    # the same loop for at least 1.2 s, still under the clock sampler: the sustained figure (the K-step region above is ~50 ms)
    sus_steps = max(args.steps, int(1.25e3 / max(ms / args.steps, 1e-3)) + 1)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s0.record(stream)
    for _ in range(sus_steps):
        step()
    s1.record(stream)
    barrier()
    sus_ms = s0.elapsed_time(s1)
    clocks = sampler.stop() if sampler else None
    ok = bool(torch.equal(x, ref))                       # forward+inverse is the identity: free correctness check

    # integer-pipe peaks of this very GPU, measured now (csrc/peaks.cu): the roofline denominators
    lo_ops, wide_ops = C.c_double(), C.c_double()
    fhe_b200.check(lib.fhe_b200_measure_int_peaks(local_rank, C.byref(lo_ops), C.byref(wide_ops), 5))
    loop_bfly = C.c_double()                                 # the butterfly alone, operands in registers: ceiling of the instruction mix
    lib.fhe_b200_measure_butterfly_loop.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_double), C.c_int]
    fhe_b200.check(lib.fhe_b200_measure_butterfly_loop(local_rank, int(chain[0]), int(chain[0]) // 3, C.byref(loop_bfly), 5))

    # per-kernel durations (second pass, events around every launch, same stream)
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
    del h

    t = torch.tensor([ms, e2e_s * 1e3, sus_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, sus_ms_max = float(t[0]), float(t[1]), float(t[2])

    # secondary metric (BASELINE.json config 4) -- every rank takes part; rank 0 reports
    hm, hm_ls = None, {}
    if not args.no_hmult:
        del x, ref
        torch.cuda.empty_cache()
        try:
            from bench_hmult import run_hmult, run_hmult_limb_sharded
            if world == 1:
                hm = run_hmult(args, local_rank, batch=8)
            else:
                hm = run_hmult(args, local_rank, batch=8, steps=5, dist=dist, lite=True)
                for b in (1, 4):
                    try:
                        hm_ls[f"batch{b}"] = run_hmult_limb_sharded(local_rank, "c4", b, 20, dist=dist)
                    except Exception as ex:
                        hm_ls[f"batch{b}"] = {"error": repr(ex)}
        except Exception as ex:  # secondary metric must never take the headline line down
            hm = {"error": repr(ex)}

    if rank == 0:
        peak, peak_src = measured_peaks()
        value = world * UNITS_PER_STEP * args.steps / (ms_max / 1e3)
        sus_value = world * UNITS_PER_STEP * sus_steps / (sus_ms_max / 1e3)
        e2e_val = world * UNITS_PER_STEP * e2e_steps / (e2e_ms_max / 1e3)
        names = {"tile_fwd": "bal_b_kernel<8,16,near,fwd> (tile-pair pass)", "tile_inv": "bal_b_kernel<8,16,near,inv> (tile-pair pass)",
                 "row_fwd": "bal_a_kernel<8,16,near,fwd> (column pass)", "row_inv": "bal_a_kernel<8,16,near,inv> (column pass)"}
        if os.environ.get("FHE_B200_NTT_BAL", "1") == "0":
            names = {"tile_fwd": "ntt_tile_fwd_kernel<12,4,16>", "tile_inv": "ntt_tile_inv_kernel<12,4,16>",
                     "row_fwd": "ntt_row_fwd_kernel<12,4,16>", "row_inv": "ntt_row_inv_kernel<12,4,16>"}
        dom = max(kinds, key=lambda k: kinds[k][1])
        tl, tms, tu = kinds[dom]
        all_ms = sum(v[1] for v in kinds.values())
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch", {}).get(dom)
            except Exception:
                traffic = None
        # ---- roofline.  The butterfly is bound by the FMA-heavy (IMAD) pipe (ncu: sm__pipe_fmaheavy_cycles_active 70-78 %, DRAM 38-47 %,
        # profiles/r01_bal_ncu_summary.md), so the binding roofline is the integer one.  Algorithmic work of one butterfly = the IMAD-class
        # instructions of the exact 64-bit Shoup multiply used: 5 IMAD.WIDE (32x32->64) + 4 IMAD (32x32->32) = 9 per lane
        # (csrc/modarith.cuh shoup_mul_lazy3; SURVEY 8d counts 10 for the canonical form).  A limb-transform is (N/2) log2 N butterflies.
        # peak = the rate at which this chip retires that 5:4 mix, from the two rates measured a moment ago:
        #     peak = 9 / (5 / R_wide + 4 / R_lo)   IMAD-class thread-operations per second,
        # achieved = butterflies/s * 9 for the dominant kernel (its own launch time); frac = achieved / peak.
        r_lo, r_wide = lo_ops.value, wide_ops.value
        bfly_per_unit = (N // 2) * LOGN
        mix_peak = 9.0 / (5.0 / r_wide + 4.0 / r_lo)
        passes_bfly = bfly_per_unit / 2.0                  # a launch covers half of the stages (8 of 16) of every limb-transform
        dom_ops = (tu * passes_bfly * 9.0) / (tms / 1e3) if tms > 0 else None
        step_ops = value / world * bfly_per_unit * 9.0
        achieved_gbs = (tu * UNIT_BYTES / 1e9) / (tms / 1e3) if tms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": config_dict(world),
            "roundtrip_bit_exact": ok and e2e_ok,
            "sustained": {"value": sus_value, "unit": UNIT, "steps": sus_steps, "seconds": sus_ms_max / 1e3,
                          "note": "same loop run for >= 1.2 s right after the K timed steps, clocks sampled across both"},
            "clocks": clocks,
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": POLYS * LIMBS * N * 8,
                    "d2h_bytes_per_step": POLYS * LIMBS * N * 8, "steps": e2e_steps,
                    "api": "fhe_b200_ntt_host (pinned host buffer, 3-stream chunk pipeline)",
                    "pcie_GBs_per_direction_per_gpu": e2e_val / world * (UNIT_BYTES / 4) / 1e9,    # a step moves every limb once in and once out for TWO transforms
                    "note": "bound by the host link: every limb is copied in (512 KiB), transformed forward and back, and copied out (512 KiB); ranks of one node "
                            "share the host's memory and root complexes, so this figure does not scale with the GPU count"},
            "roofline": {"bound": "imad", "achieved": (dom_ops / 1e12) if dom_ops else None, "peak": mix_peak / 1e12, "unit": "T IMAD-class op/s",
                         "frac": (dom_ops / mix_peak) if dom_ops else None, "traffic": traffic,
                         "kernel": names[dom], "launches_profiled": tl, "avg_launch_ms": (tms / tl) if tl else None,
                         "units_per_launch": (tu / tl) if tl else None,
                         "kernel_share_of_step": (tms / all_ms) if all_ms else None,
                         "ops_per_butterfly": {"IMAD.WIDE": 5, "IMAD": 4}, "butterflies_per_limb_transform": bfly_per_unit,
                         "measured_now": {"imad_lo_Tops": r_lo / 1e12, "imad_wide_Tops": r_wide / 1e12,
                                          "how": "fhe_b200_measure_int_peaks: 8 independent chains per thread, 64 warps per SM, best of 5"},
                         "peak_formula": "9 / (5 / imad_wide + 4 / imad_lo)",
                         "register_only_butterfly_loop": {
                             "Gbutterflies_s": loop_bfly.value / 1e9, "frac_of_peak": loop_bfly.value * 9 / mix_peak,
                             "whole_step_frac_of_loop": (step_ops / 9.0) / loop_bfly.value if loop_bfly.value else None,
                             "how": "fhe_b200_measure_butterfly_loop: 4 forward stages on 16 register-resident values per thread, same multiply, bounds "
                                    "and reductions as the passes, no memory traffic; what the carry / butterfly additions cost on top of the multiplier"},
                         "whole_step": {"achieved": step_ops / 1e12, "frac": step_ops / mix_peak,
                                        "peak_limb_transforms_s": mix_peak / 9.0 / bfly_per_unit},
                         "survey_8d_formula": {"imad_ops_per_limb_transform": 10 * bfly_per_unit,
                                               "frac_vs_imad_lo_peak": value / world * 10 * bfly_per_unit / r_lo,
                                               "frac_vs_imad_wide_peak": value / world * 10 * bfly_per_unit / r_wide},
                         "per_kernel_ms": {k: v[1] / psteps for k, v in kinds.items()}},
            # HBM view of the same numbers (not the binding roofline): algorithmic 1 MiB per limb-transform for the whole step;
            # each of the two passes of a transform moves that 1 MiB once (ncu dram bytes per launch = `traffic`)
            "hbm": {"peak": peak, "unit": "GB/s", "peak_source": peak_src, "bytes_per_unit": UNIT_BYTES,
                    "step_achieved": value * UNIT_BYTES / 1e9 / world, "step_frac": value * UNIT_BYTES / 1e9 / world / peak,
                    "dominant_kernel_achieved": achieved_gbs, "dominant_kernel_frac": (achieved_gbs / peak) if achieved_gbs else None,
                    "wasted_traffic_vs_algorithmic": 2.0},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample()
        if hm is not None:
            line["hmult"] = hm
        if hm_ls:
            line["hmult_limb_sharded"] = hm_ls
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
