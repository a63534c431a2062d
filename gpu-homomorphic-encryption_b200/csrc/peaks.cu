// Measured integer-pipe peaks of the device the benchmark runs on (the roofline denominators of bench.py).
//
// The NTT butterfly is bound by the FMA-heavy (IMAD) pipe, not by HBM, and MEASURED_PEAKS.json carries no
// integer figure, so the benchmark measures the two instruction classes the butterfly is made of in the same run:
//   IMAD (32 x 32 -> low 32, mad.lo.u32)  and  IMAD.WIDE (32 x 32 -> 64, mul.wide.u32),
// each as 8 independent dependency chains per thread, 64 resident warps per SM, operands that change every iteration (a
// loop-invariant product would be hoisted).  Result: warp-wide operations per second over the whole chip.
#include "common.cuh"

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

}  // namespace fhe_b200

using namespace fhe_b200;

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
