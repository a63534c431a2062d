// Measured integer-pipe peaks of the device the benchmark runs on (the roofline denominators of bench.py).
//
// The NTT butterfly is bound by the FMA-heavy (IMAD) pipe, not by HBM, and MEASURED_PEAKS.json carries no
// integer figure, so the benchmark measures the two instruction classes the butterfly is made of in the same run:
//   IMAD (32 x 32 -> low 32, mad.lo.u32)  and  IMAD.WIDE (32 x 32 -> 64, mul.wide.u32),
// each as 8 independent dependency chains per thread, 64 resident warps per SM, operands that change every iteration (a
// loop-invariant product would be hoisted).  Result: warp-wide operations per second over the whole chip.
#include "common.cuh"
#include "ntt_core.cuh"

namespace fhe_b200 {

constexpr int kPeakIters = 2048, kPeakIlp = 8;

__global__ void __launch_bounds__(256) peak_imad_lo_kernel(u32* out, u32 a, u32 b) {
    u32 x[kPeakIlp];
#pragma unroll
    for (int i = 0; i < kPeakIlp; i++) x[i] = threadIdx.x * 7 + i;
    for (int it = 0; it < kPeakIters; it++) {
#pragma unroll
        for (int i = 0; i < kPeakIlp; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < kPeakIlp; i++) s ^= x[i];
    if (s == 0x12345678u) out[0] = s;
}
__global__ void __launch_bounds__(256) peak_imad_wide_kernel(u64* out, u32 a) {
    u64 x[kPeakIlp];
#pragma unroll
    for (int i = 0; i < kPeakIlp; i++) x[i] = threadIdx.x * 7 + i;
    for (int it = 0; it < kPeakIters; it++) {
#pragma unroll
        for (int i = 0; i < kPeakIlp; i++) {
            const u32 lo = (u32)x[i] ^ (u32)(x[i] >> 32);
            asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(x[i]) : "r"(lo), "r"(a));
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < kPeakIlp; i++) s ^= x[i];
    if (s == 0x12345678u) out[0] = s;
}

// The butterfly itself with everything else removed: four forward stages (32 lazy Harvey butterflies with the engine's own
// shoup_mul_lazy3, bounds and range reductions: fwd_stages of ntt_core.cuh) on sixteen values that never leave the registers, one
// twiddle pair per stage from registers, 256-thread CTAs, as many resident as the 64 registers allow.  What this loop reaches is the
// practical ceiling of the INSTRUCTION MIX (IMAD.WIDE + IMAD + the carry and butterfly additions that cannot all hide under the
// multiplier): the passes' distance from it is what loads, stores, exchanges and twiddle traffic cost.
struct TwRegs {
    Twiddle t[4];
    FHE_HD Twiddle get(int v, int) const { return t[v]; }
};
constexpr int kMixIters = 256;
__global__ void __launch_bounds__(256, 4) peak_butterfly_kernel(u64* out, LimbParams P, Twiddle t0, Twiddle t1, Twiddle t2, Twiddle t3) {
    u64 x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (u64)(threadIdx.x * 16 + i) * 0x9E3779B97F4A7C15ULL % P.q;
    TwRegs tw{{t0, t1, t2, t3}};
#pragma unroll 1
    for (int it = 0; it < kMixIters; it++) {
        fwd_stages<4, 4, 16, true, 1>(x, tw, P);
        // back under 2q for the next round (the cost of one range reduction per value per four stages is part of the real passes too)
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = near60_reduce(x[i], 0 - P.q);
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= x[i];
    if (s == 0x12345678u) out[0] = s;
}

}  // namespace fhe_b200

using namespace fhe_b200;

// lazy butterflies per second of the register-only loop above (whole chip), best of `reps` launches; q must be a chain prime
// (2^60 - 2^32 < q < 2^60), w any residue
extern "C" int fhe_b200_measure_butterfly_loop(int device, uint64_t q, uint64_t w, double* butterflies_per_s, int reps) {
    FHE_REQUIRE(butterflies_per_s && (q >> 60) == 0 && q > (1ull << 60) - (1ull << 32) && w < q, "measure_butterfly_loop: bad argument");
    if (reps < 1) reps = 5;
    DeviceGuard dev_guard(device);
    cudaDeviceProp prop;
    FHE_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 4 * 4;            // four waves of four resident CTAs
    void* buf = nullptr;
    FHE_CUDA(cudaMalloc(&buf, 256));
    LimbParams P{};
    P.q = q; P.tq = kTQ * q;
    Twiddle t[4];
    for (int i = 0; i < 4; i++) {
        const uint64_t wi = (uint64_t)(((unsigned __int128)w * (2 * i + 3)) % q);
        t[i].w = wi; t[i].ws = (uint64_t)((((unsigned __int128)wi) << 64) / q);
    }
    cudaEvent_t e0, e1;
    FHE_CUDA(cudaEventCreate(&e0)); FHE_CUDA(cudaEventCreate(&e1));
    const double bfly = (double)blocks * 256 * kMixIters * 32;
    double best = 0;
    for (int r = 0; r < reps + 1; r++) {
        cudaEventRecord(e0, 0);
        peak_butterfly_kernel<<<blocks, 256>>>((u64*)buf, P, t[0], t[1], t[2], t[3]);
        count_launch();
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms > 0 && bfly / (ms * 1e-3) > best) best = bfly / (ms * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    FHE_CUDA(cudaGetLastError());
    *butterflies_per_s = best;
    return 0;
}

// ops/s (thread-level 32-bit operations) of IMAD.lo and IMAD.WIDE, best of `reps` timed launches each, on `stream`'s device
extern "C" int fhe_b200_measure_int_peaks(int device, double* imad_lo_ops, double* imad_wide_ops, int reps) {
    FHE_REQUIRE(imad_lo_ops && imad_wide_ops, "measure_int_peaks: null argument");
    if (reps < 1) reps = 5;
    DeviceGuard dev_guard(device);
    cudaDeviceProp prop;
    FHE_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8;
    void* buf = nullptr;
    FHE_CUDA(cudaMalloc(&buf, 256));
    cudaEvent_t e0, e1;
    FHE_CUDA(cudaEventCreate(&e0)); FHE_CUDA(cudaEventCreate(&e1));
    const double ops = (double)blocks * 256 * kPeakIters * kPeakIlp;
    double best[2] = {0, 0};
    for (int k = 0; k < 2; k++)
        for (int r = 0; r < reps + 1; r++) {
            cudaEventRecord(e0, 0);
            if (k == 0) peak_imad_lo_kernel<<<blocks, 256>>>((u32*)buf, 0x9e3779b9u, 0x7f4a7c15u);
            else peak_imad_wide_kernel<<<blocks, 256>>>((u64*)buf, 0x9e3779b9u);
            count_launch();
            cudaEventRecord(e1, 0);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0 && ms > 0 && ops / (ms * 1e-3) > best[k]) best[k] = ops / (ms * 1e-3);
        }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    FHE_CUDA(cudaGetLastError());
    *imad_lo_ops = best[0]; *imad_wide_ops = best[1];
    return 0;
}
