"""profiles/r01_lincomb_mma_ncu_summary.md from two `ncu --page raw --csv` dumps of the eleven conversions of one HMult
(tensor-core kernels, IMAD kernels):  python tools/lincomb_profile_summary.py mma_raw.csv imad_raw.csv"""
import csv, sys


def load(p):
    rows = list(csv.reader(open(p))); hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}; units = rows[1]
    out = []
    for r in rows[2:]:
        g = lambda k: r[idx[k]] if k in idx else ''
        st = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(r[i] or 0)
              for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h}
        tot = sum(st.values()); top = ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
        t = float(g("gpu__time_duration.sum")); u = units[idx["gpu__time_duration.sum"]]
        us = t * 1000 if u in ("ms", "msecond") else t / 1000 if u in ("ns", "nsecond") else t
        out.append((g("Kernel Name").replace("void ", "").replace("(LcMmaArgs)", "").replace("(LcKernelArgs)", ""), us,
                    float(g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or 0),
                    float(g("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed")),
                    float(g("smsp__issue_active.avg.pct_of_peak_sustained_active")), top))
    return out


mma, imad = load(sys.argv[1]), load(sys.argv[2])
names_mma = ["Q->R (24 -> 25 limbs), all four inputs in one launch", "scale-and-round t/Q (24 (+1 extra) -> 25), 3 polynomials", "R->Q (25 -> 24), 3 polynomials",
             "ModUp digit 0 (8 -> 24)", "ModUp digit 1", "ModUp digit 2", "ModDown (8 -> 24, fused epilogue), c0", "ModDown, c1"]
print("# r01: tensor-core vs IMAD base conversion, `ncu --set full --clock-control none` of the conversions of one BFV")
print("# multiply+relinearize at config 4, batch 4 (`python tools/prof_hmult.py 4`; IMAD run: FHE_B200_LINCOMB_MMA=0).  Same process, same inputs,")
print("# results bit-identical (tests/test_gpu_bfv.py).  Durations are ncu's serialised cold-cache times: compare the two columns.")
print()
print("| conversion | IMAD kernel | us | FMA-heavy % | tensor-core kernel | us | tensor % | FMA-heavy % | issue % | top stalls (tensor-core kernel) |")
print("|---|---|---:|---:|---|---:|---:|---:|---:|---|")
ti = tm = 0.0
names11 = ["Q->R (24 -> 25 limbs), input 0", "Q->R, input 1", "Q->R, input 2", "Q->R, input 3"] + names_mma[1:]
if len(mma) == 11:
    names_mma = names11
n = min(len(mma), len(imad))
for i in range(n):
    a, b = imad[i], mma[i]
    ti += a[1]; tm += b[1]
    nm = names_mma[i] if len(mma) == len(names_mma) and i < len(names_mma) else f"conversion {i}"
    print(f"| {nm} | `{a[0]}` | {a[1]:.1f} | {a[3]:.0f} | `{b[0]}` | {b[1]:.1f} | {b[2]:.0f} | {b[3]:.0f} | {b[4]:.0f} | {b[5]} |")
print(f"| **all** | | **{ti:.0f}** | | | **{tm:.0f}** | | | | |")
print()
print("Reading: the int8-decomposed GEMM (`mma.sync.m16n8k32.u8.u8.s32`, Toeplitz byte matrix, 128 byte-MACs per 64-bit MAC) takes the")
print("multiply-accumulates off the FMA-heavy pipe, which the IMAD kernel keeps 54-67% busy with four IMAD.WIDE per (source, target) pair.")
print("What remains on that pipe is the prologue (x_i * (Q/q_i)^-1 mod q_i, the 128-bit fixed-point overflow count) and the epilogue (Barrett")
print("128 -> 64, ModDown), now the larger share of the kernel; `wait` (fixed-latency dependencies at 16 resident warps per SM) is the top")
print(f"stall.  Net: all conversions of one multiply {ti:.0f} us -> {tm:.0f} us (x{ti / tm:.2f}).  CUDA-event timing inside `bench_hmult.py` (warm, back to")
print("back) at batch 4: 2.21 ms (IMAD) -> 1.40 ms per four multiplies; HMult+relinearize 995 -> 1295 ops/s.  Kept as the default for S*T >= 64.")
