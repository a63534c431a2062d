"""profiles/r02_sweep_c5.md from gpurun_out/sweep_c5_{1,2,4,8}.jsonl (tools/sweep_c5.py)."""
import json, os, sys
rows = {}
for g in (1, 2, 4, 8):
    p = f"gpurun_out/sweep_c5_{g}.jsonl"
    if os.path.exists(p):
        rows[g] = [json.loads(l) for l in open(p)]
one = rows[1]
pk = [r for r in one if r["what"] == "peaks"][0]
out = ["# r02: BASELINE.json config 5 -- throughput sweep N = 2^12..2^17 x batch 1..1024 ciphertexts at 1 / 2 / 4 / 8 B200 (`tools/sweep_c5.py`)",
       "",
       f"One GPU, CUDA events, every point checked (round trip bit-exact; decrypt(HMult) == product).  Integer peaks measured in the same run: IMAD {pk['imad_lo_Tops']:.2f} T/s, "
       f"IMAD.WIDE {pk['imad_wide_Tops']:.2f} T/s -> butterfly-mix peak {pk['mix_peak_Tops']:.2f} T IMAD-class op/s (bench.py's roofline).  CPU column: the oracle's Harvey/Shoup NTT, "
       "gcc -O3 -march=native, all host threads of the GPU box.  A ciphertext = 2 polynomials of L(N) limbs.", "",
       "## M1: forward + inverse RNS-NTT (limb-transforms/s), one GPU", "",
       "| N | L | ciphertexts | limb-transforms/s | IMAD-mix roofline frac | CPU port (threads) | x CPU |", "|---|---:|---:|---:|---:|---:|---:|"]
for r in one:
    if r["what"] == "M1":
        out.append(f"| 2^{r['logn']} | {r['limbs']} | {r['ciphertexts_per_gpu']} | {r['limb_transforms_per_s']:,} | {r['imad_mix_roofline_frac']:.3f} | "
                   f"{r['cpu_port_limb_transforms_per_s']:,} ({r['cpu_threads']}) | {r['speedup_vs_cpu']} |")
out += ["", "## M2: BFV HMult + relinearize (ops/s), one GPU", "",
        "The reference column is empty by construction: `relinearize` is `components.resize(2)` (src/fhe.cu:226-235), `multiply` never scales by t/q, and its NTT launches are illegal for N > 1024 (src/ntt.cu:36-39).", "",
        "| N | L | R | K | dnum | ciphertext pairs | ms per op | ops/s | decrypts to the product |", "|---|---:|---:|---:|---:|---:|---:|---:|---|"]
for r in one:
    if r["what"] == "M2":
        out.append(f"| 2^{r['logn']} | {r['L']} | {r['R']} | {r['K']} | {r['dnum']} | {r['ciphertext_pairs_per_gpu']} | {r['ms_per_op']} | {r['hmult_relin_ops_per_s']:,} | {r['decrypts_to_product']} |")
out += ["", "## 1 / 2 / 4 / 8 GPUs (batch sharding: every rank its own ciphertexts, no data-path collective; time = max over ranks)", "",
        "| metric | N | per-GPU batch | " + " | ".join(f"{g} GPU" for g in sorted(rows)) + " | x at " + str(max(rows)) + " |", "|---|---|---:|" + "---:|" * (len(rows) + 1)]
def val(g, what, logn, batch):
    for r in rows[g]:
        if r["what"] == what and r["logn"] == logn and r.get("ciphertexts_per_gpu", r.get("ciphertext_pairs_per_gpu")) == batch:
            return r.get("limb_transforms_per_s", r.get("hmult_relin_ops_per_s"))
    return None
for what, name in (("M1", "RNS-NTT limb-transforms/s"), ("M2", "HMult+relin ops/s")):
    for logn in (13, 16):
        for batch in (1, 16, 256):
            v = {g: val(g, what, logn, batch) for g in sorted(rows)}
            if v[1] is None or any(v[g] is None for g in v):
                continue
            out.append(f"| {name} | 2^{logn} | {batch} | " + " | ".join(f"{v[g]:,.0f}" for g in sorted(rows)) + f" | {v[max(rows)] / v[1]:.2f} |")
out += ["", "The limb-sharded multiply (one batch on all GPUs, peer stores over NVLink) is in profiles/r02_sharded_hmult.md and in bench.py's `hmult_limb_sharded`.", "",
        "## Reference CUDA kernels on the same box (compiled from /root/reference into oracle/_ref, `oracle/ref_ntt_kernels.cu`)", "",
        "The reference launches one polynomial per call on 256-bit elements (`ntt_pointwise_mul_kernel`, kernels/ntt_kernels.cu:124-137 via src/ntt.cu:65-68; `poly_add_kernel`, "
        "src/polynomial.cu:70-82).  Times are microseconds per polynomial of N coefficients; the engine's are for 64-bit residues, per polynomial inside one launch of 64 "
        "(the one-launch-per-polynomial figure is dominated by the Python/ctypes call, not by the device).", "",
        "| N | ref pointwise mul (256-bit Montgomery) | ref poly_add | ours poly_mul, batched | ours poly_add, batched | ours poly_mul, one call per polynomial |", "|---|---:|---:|---:|---:|---:|"]
for r in one:
    if r["what"] == "reference_kernels":
        out.append(f"| 2^{r['logn']} | {r['ref_ntt_pointwise_mul_us']:.2f} | {r['ref_poly_add_us']:.2f} | {r['ours_poly_mul_us_per_poly_batched64']:.3f} | {r['ours_poly_add_us_per_poly_batched64']:.3f} | {r['ours_poly_mul_us_one_launch_per_poly']:.1f} |")
n = [r for r in one if r["what"] == "reference_ntt_forward_n1024"][0]
out += ["", f"`NTTEngine::forward` at N = 1024 (the only size the reference can launch: one block of N threads, src/ntt.cu:30-40; placeholder twiddles, output meaningless, only timed): "
        f"**{n['ref_us_per_transform']:.1f} us** per transform against {n['ours_us_one_transform']:.1f} us for one transform here (launch-bound) and "
        f"{n['ours_us_per_transform_batched4096'] * 1e3:.1f} ns per transform in a batch of 4096.  Everything above N = 1024 and all of HMult+relinearize: reference launch fails / not implemented."]
open("profiles/r02_sweep_c5.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[-14:]))
