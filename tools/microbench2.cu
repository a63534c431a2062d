// Per-instruction throughput on the integer pipes (ops per clock per SM), and candidate butterfly formulations.
// Each kernel runs ILP independent chains per thread, 64 warps per SM.  Prints one JSON object per line.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cstdlib>

constexpr int ITERS = 2048;
constexpr int ILP = 8;
typedef uint64_t u64;
typedef uint32_t u32;

#define KERNEL32(NAME, BODY)                                                        \
    __global__ void NAME(u32* out, u32 a, u32 b) {                                  \
        u32 x[ILP];                                                                 \
        _Pragma("unroll") for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 7 + i; \
        for (int it = 0; it < ITERS; it++) {                                        \
            _Pragma("unroll") for (int i = 0; i < ILP; i++) { BODY; }               \
        }                                                                           \
        u32 s = 0;                                                                  \
        _Pragma("unroll") for (int i = 0; i < ILP; i++) s ^= x[i];                  \
        if (s == 0x12345678u) out[0] = s;                                           \
    }
#define KERNEL64(NAME, BODY)                                                        \
    __global__ void NAME(u64* out, u32 a, u32 b) {                                  \
        u64 x[ILP];                                                                 \
        _Pragma("unroll") for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 7 + i; \
        for (int it = 0; it < ITERS; it++) {                                        \
            _Pragma("unroll") for (int i = 0; i < ILP; i++) { BODY; }               \
        }                                                                           \
        u64 s = 0;                                                                  \
        _Pragma("unroll") for (int i = 0; i < ILP; i++) s ^= x[i];                  \
        if (s == 0x12345678u) out[0] = s;                                           \
    }

KERNEL32(k_mad_lo, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b)))
KERNEL32(k_mul_lo, asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a)))
KERNEL32(k_mad_hi, asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b)))
KERNEL32(k_mul_hi, asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a)))
KERNEL32(k_add, asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a)))
KERNEL32(k_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b)))
KERNEL32(k_shf, asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b)))
KERNEL32(k_setp_selp, asm volatile("{ .reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(x[i]) : "r"(a), "r"(b)))
KERNEL32(k_mix_lo_add, asm volatile("{ .reg .u32 t; mad.lo.u32 t, %0, %1, %2; add.u32 %0, t, %1; }" : "+r"(x[i]) : "r"(a), "r"(b)))
KERNEL32(k_mix_lo_2add, asm volatile("{ .reg .u32 t; mad.lo.u32 t, %0, %1, %2; add.u32 t, t, %1; xor.b32 %0, t, %2; }" : "+r"(x[i]) : "r"(a), "r"(b)))
KERNEL64(k_mul_wide, { u32 lo = (u32)x[i] ^ (u32)(x[i] >> 32); asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(x[i]) : "r"(lo), "r"(a)); })
KERNEL64(k_mad_wide, { u32 lo = (u32)x[i]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(lo), "r"(a)); })
KERNEL64(k_mad_wide_indep, { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(a), "r"(b)); })
KERNEL64(k_mix_wide_add, { u32 lo = (u32)x[i]; asm volatile("{ .reg .u64 t; mul.wide.u32 t, %1, %2; add.u64 %0, t, %0; }" : "+l"(x[i]) : "r"(lo), "r"(a)); })
KERNEL64(k_add64, asm volatile("add.u64 %0, %0, %1;" : "+l"(x[i]) : "l"((u64)a << 20 | b)))
KERNEL64(k_mul_lo64, x[i] = x[i] * (((u64)a << 32) | b))
KERNEL64(k_mul_hi64, x[i] = __umul64hi(x[i], (((u64)a << 32) | b)) + i)
KERNEL64(k_csub64, { const u64 m = ((u64)a << 32) | b; x[i] = x[i] + 12345; x[i] = x[i] >= m ? x[i] - m : x[i]; })
KERNEL64(k_dfma, { double d = __longlong_as_double(x[i]); d = fma(d, 1.0000001, 0.5); x[i] = __double_as_longlong(d); })

// ---- butterflies (registers only): X' = X + T, Y' = X + 2q - T, T = Shoup(Y, w) ----
__device__ __forceinline__ u64 shoup_std(u64 y, u64 w, u64 ws, u64 q) { return y * w - __umul64hi(y, ws) * q; }
// approximate quotient from three 32x32 products: h' in [h-2, h]  ->  T in [0, 4q)
__device__ __forceinline__ u64 shoup_approx(u64 y, u64 w, u64 ws, u64 nq) {
    const u32 y0 = (u32)y, y1 = (u32)(y >> 32), s0 = (u32)ws, s1 = (u32)(ws >> 32);
    const u64 cross = (u64)__umulhi(y1, s0) + (u64)__umulhi(y0, s1);
    const u64 h = (u64)y1 * s1 + cross;
    return y * w + h * nq;            // nq = -q mod 2^64
}
template <int MODE>
__global__ void k_bfly(u64* out, u64 w, u64 ws, u64 q) {
    u64 x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 3 + i; }
    const u64 twoq = 2 * q, fourq = 4 * q, nq = 0 - q, eightq = 8 * q;
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            u64 X = x[i];
            if (MODE == 0) {            // Harvey: csub every stage
                X = X >= twoq ? X - twoq : X;
                const u64 t = shoup_std(y[i], w, ws, q);
                x[i] = X + t; y[i] = X + twoq - t;
            } else if (MODE == 1) {     // standard Shoup, no csub (lazy head-room)
                const u64 t = shoup_std(y[i], w, ws, q);
                x[i] = X + t; y[i] = X + twoq - t;
            } else if (MODE == 2) {     // approximate quotient, no csub
                const u64 t = shoup_approx(y[i], w, ws, nq);
                x[i] = X + t; y[i] = X + fourq - t;
            } else {                    // approximate quotient + csub(8q) every stage
                X = X >= eightq ? X - eightq : X;
                const u64 t = shoup_approx(y[i], w, ws, nq);
                x[i] = X + t; y[i] = X + fourq - t;
            }
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i] ^ y[i];
    if (s == 0x12345678) out[0] = s;
}

template <class F> static float time_ms(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    void* buf; cudaMalloc(&buf, 1 << 20);
    const int bps = argc > 1 ? atoi(argv[1]) : 8;     // resident 256-thread blocks per SM (occupancy sweep)
    const int blocks = sms * bps, threads = 256;
    const double ghz = p.clockRate / 1e6;
    auto report = [&](const char* name, float ms, double ops_per_iter) {
        const double ops = (double)blocks * threads * ITERS * ILP * ops_per_iter;
        printf("{\"name\": \"%s\", \"ms\": %.4f, \"Tops\": %.3f, \"ops_per_clk_per_sm\": %.1f}\n", name, ms, ops / ms / 1e9,
               ops / (ms * 1e-3) / (ghz * 1e9) / sms);
    };
#define RUN32(K, N) report(#K, time_ms([&] { K<<<blocks, threads>>>((u32*)buf, 0x9e3779b9u, 0x7f4a7c15u); }), N)
#define RUN64(K, N) report(#K, time_ms([&] { K<<<blocks, threads>>>((u64*)buf, 0x9e3779b9u, 0x7f4a7c15u); }), N)
    RUN32(k_mad_lo, 1); RUN32(k_mul_lo, 1); RUN32(k_mad_hi, 1); RUN32(k_mul_hi, 1); RUN32(k_add, 1); RUN32(k_lop3, 1); RUN32(k_shf, 1);
    RUN32(k_setp_selp, 2); RUN32(k_mix_lo_add, 2); RUN32(k_mix_lo_2add, 3);
    RUN64(k_mul_wide, 1); RUN64(k_mad_wide, 1); RUN64(k_mad_wide_indep, 1); RUN64(k_mix_wide_add, 1); RUN64(k_add64, 1);
    RUN64(k_mul_lo64, 1); RUN64(k_mul_hi64, 1); RUN64(k_csub64, 1); RUN64(k_dfma, 1);
    const u64 q = 0xffffffffffc0001ULL, w = 0x123456789abcdefULL, ws = (u64)(((unsigned __int128)w << 64) / q);
    const char* names[4] = {"bfly_harvey", "bfly_lazy", "bfly_approx_lazy", "bfly_approx_csub"};
    float t;
    t = time_ms([&] { k_bfly<0><<<blocks, threads>>>((u64*)buf, w, ws, q); }); printf("{\"name\": \"%s\", \"Gbfly_per_s\": %.1f}\n", names[0], (double)blocks * threads * (ITERS / 4) * ILP / t / 1e6);
    t = time_ms([&] { k_bfly<1><<<blocks, threads>>>((u64*)buf, w, ws, q); }); printf("{\"name\": \"%s\", \"Gbfly_per_s\": %.1f}\n", names[1], (double)blocks * threads * (ITERS / 4) * ILP / t / 1e6);
    t = time_ms([&] { k_bfly<2><<<blocks, threads>>>((u64*)buf, w, ws, q); }); printf("{\"name\": \"%s\", \"Gbfly_per_s\": %.1f}\n", names[2], (double)blocks * threads * (ITERS / 4) * ILP / t / 1e6);
    t = time_ms([&] { k_bfly<3><<<blocks, threads>>>((u64*)buf, w, ws, q); }); printf("{\"name\": \"%s\", \"Gbfly_per_s\": %.1f}\n", names[3], (double)blocks * threads * (ITERS / 4) * ILP / t / 1e6);
    return 0;
}
