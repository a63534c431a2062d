// Measures the integer-pipe and copy ceilings the NTT roofline is quoted against (BASELINE.md section 2:
// "32-bit IMAD issue rate: not yet measured").  Dependent-free chains, 8 independent accumulators per thread.
// Prints one JSON object.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/microbench.cu -o tools/microbench.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int ILP = 8;

__global__ void k_imad(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    if (s == 0x12345678) out[0] = s;
}
__global__ void k_imad_wide(uint64_t* out, uint32_t a, uint32_t b) {
    uint64_t acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            uint32_t lo = (uint32_t)acc[i];
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(lo), "r"(a));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    if (s == 0x12345678 + b) out[0] = s;
}
__global__ void k_iadd3(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; xor.b32 %0, t, %2; }" : "+r"(acc[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    if (s == 0x12345678) out[0] = s;
}
// 64-bit Shoup butterfly throughput, registers only (the NTT inner loop without memory)
__global__ void k_shoup(uint64_t* out, uint64_t w, uint64_t ws, uint64_t q) {
    uint64_t x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 3 + i; }
    const uint64_t twoq = 2 * q;
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            uint64_t h = __umul64hi(y[i], ws);
            uint64_t t = y[i] * w - h * q;
            uint64_t X = x[i];
            X = X >= twoq ? X - twoq : X;
            x[i] = X + t; y[i] = X + twoq - t;
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i] ^ y[i];
    if (s == 0x12345678) out[0] = s;
}
__global__ void k_copy(ulonglong2* __restrict__ dst, const ulonglong2* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

template <class F> static float time_ms(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    void* buf; CK(cudaMalloc(&buf, 1 << 20));
    const int blocks = sms * 8, threads = 256;
    const double lanes = (double)blocks * threads;
    float t1 = time_ms([&] { k_imad<<<blocks, threads>>>((uint32_t*)buf, 3, 5); }, 5);
    float t2 = time_ms([&] { k_imad_wide<<<blocks, threads>>>((uint64_t*)buf, 3, 5); }, 5);
    float t3 = time_ms([&] { k_iadd3<<<blocks, threads>>>((uint32_t*)buf, 3, 5); }, 5);
    float t4 = time_ms([&] { k_shoup<<<blocks, threads>>>((uint64_t*)buf, 0x123456789abcdefULL, 0x23456789abcdef01ULL, 0xffffffffffc0001ULL); }, 5);
    const size_t nbytes = (size_t)2 << 30;
    void *s, *d; CK(cudaMalloc(&s, nbytes)); CK(cudaMalloc(&d, nbytes));
    CK(cudaMemset(s, 1, nbytes));
    float t5 = time_ms([&] { k_copy<<<sms * 16, 256>>>((ulonglong2*)d, (const ulonglong2*)s, nbytes / 16); }, 5);
    float t6 = time_ms([&] { cudaMemcpyAsync(d, s, nbytes, cudaMemcpyDeviceToDevice); }, 5);
    const double ops = lanes * ITERS * ILP;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, "
           "\"imad_lo_Tops\": %.3f, \"imad_wide_Tops\": %.3f, \"alu_pair_Tops\": %.3f, \"shoup_butterfly_G_per_s\": %.2f, "
           "\"copy_kernel_GBs\": %.1f, \"memcpy_d2d_GBs\": %.1f}\n",
           p.name, sms, p.clockRate, ops / t1 / 1e9, ops / t2 / 1e9, 2 * ops / t3 / 1e9,
           lanes * (ITERS / 4) * ILP / t4 / 1e6, 2.0 * nbytes / t5 / 1e6, 2.0 * nbytes / t6 / 1e6);
    return 0;
}
