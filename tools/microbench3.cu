// Ceiling experiment for the balanced NTT design: a radix-16 register round (4 forward stages, 32 butterflies per thread) on
// shared memory with one block barrier per round -- the inner loop of a pass with two rounds.  Variants: butterfly formulation
// (V), range-reduction flavour (RED, applied to X in one stage of four), resident CTAs per SM.  Prints G butterflies/s.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
struct alignas(16) Tw { u64 w, ws; };
#define D __device__ __forceinline__

D u64 shoup4_asm(u64 x, u64 w, u64 ws, u64 nq) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 y0, y1, s0, s1, w0, w1, n0, n1, a, b, h0, h1, t0, t1;\n\t"
        ".reg .u64 h, t;\n\t"
        "mov.b64 {y0, y1}, %1;\n\t" "mov.b64 {w0, w1}, %2;\n\t" "mov.b64 {s0, s1}, %3;\n\t" "mov.b64 {n0, n1}, %4;\n\t"
        "mul.hi.u32 a, y1, s0;\n\t" "mul.hi.u32 b, y0, s1;\n\t" "mov.b64 h, {a, 0};\n\t" "mad.wide.u32 h, y1, s1, h;\n\t"
        "mov.b64 {h0, h1}, h;\n\t" "add.cc.u32 h0, h0, b;\n\t" "addc.u32 h1, h1, 0;\n\t"
        "mul.wide.u32 t, y0, w0;\n\t" "mov.b64 {t0, t1}, t;\n\t" "mad.lo.u32 t1, y0, w1, t1;\n\t" "mad.lo.u32 t1, y1, w0, t1;\n\t"
        "mov.b64 t, {t0, t1};\n\t" "mad.wide.u32 t, h0, n0, t;\n\t" "mov.b64 {t0, t1}, t;\n\t"
        "mad.lo.u32 t1, h0, n1, t1;\n\t" "mad.lo.u32 t1, h1, n0, t1;\n\t" "mov.b64 %0, {t0, t1};\n\t"
        "}" : "=l"(r) : "l"(x), "l"(w), "l"(ws), "l"(nq));
    return r;
}
D u64 shoup4_c(u64 Y, u64 w, u64 ws, u64 nq) {
    const u32 y0 = (u32)Y, y1 = (u32)(Y >> 32), s0 = (u32)ws, s1 = (u32)(ws >> 32), w0 = (u32)w, w1 = (u32)(w >> 32);
    const u32 n0 = (u32)nq, n1 = (u32)(nq >> 32);
    const u32 a = __umulhi(y1, s0), b = __umulhi(y0, s1);
    const u64 h = (u64)y1 * s1 + a + b;
    const u32 h0 = (u32)h, h1 = (u32)(h >> 32);
    u64 t = (u64)y0 * w0;
    t = (u64)h0 * n0 + t;
    const u32 t1 = (u32)(t >> 32) + y0 * w1 + y1 * w0 + h0 * n1 + h1 * n0;
    return ((u64)t1 << 32) | (u32)t;
}
// near-2^60 modulus (q = 2^60 - delta, delta < 2^28): the h0*n1 term is a shift
D u64 shoup4_near(u64 Y, u64 w, u64 ws, u64 nq) {
    const u32 y0 = (u32)Y, y1 = (u32)(Y >> 32), s0 = (u32)ws, s1 = (u32)(ws >> 32), w0 = (u32)w, w1 = (u32)(w >> 32);
    const u32 n0 = (u32)nq;
    const u32 a = __umulhi(y1, s0), b = __umulhi(y0, s1);
    const u64 h = (u64)y1 * s1 + a + b;
    const u32 h0 = (u32)h, h1 = (u32)(h >> 32);
    u64 t = (u64)y0 * w0;
    t = (u64)h0 * n0 + t;
    const u32 t1 = (u32)(t >> 32) + y0 * w1 + y1 * w0 - (h0 << 28) + h1 * n0;
    return ((u64)t1 << 32) | (u32)t;
}
D u64 red_near60(u64 x, u64 nq) { return x + (x >> 60) * nq; }
D u64 red_csub(u64 x, u64 m) { return x >= m ? x - m : x; }
D u64 red_mask(u64 x, u32 delta) { return (x & 0x0fffffffffffffffULL) + (u32)((u32)(x >> 60) * delta); }

template <int V, int RED, int MINB>
__global__ void __launch_bounds__(256, MINB) k(u64* g, const Tw* gt, u64 q, int rounds) {
    __shared__ u64 s[4096];
    __shared__ Tw st[16];
    const u32 tid = threadIdx.x;
    for (int i = 0; i < 16; i++) s[tid + 256 * i] = g[(blockIdx.x & 1023) * 4096 + tid + 256 * i];
    if (tid < 16) st[tid] = gt[tid];
    __syncthreads();
    const u64 nq = 0 - q, fourq = 4 * q, eightq = 8 * q;
    const u32 delta = (u32)nq;
    u64 x[16];
#pragma unroll 1
    for (int r = 0; r < rounds; r++) {
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = s[(e << 8) | tid];
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const int stp = 8 >> v;
#pragma unroll
            for (int key = 0; key < (1 << v); key++) {
                const Tw w = st[(1 << v) - 1 + key];
#pragma unroll
                for (int j = 0; j < stp; j++) {
                    const int e = (key << (4 - v)) | j;
                    u64 X = x[e];
                    if (v == 1) {
                        if (RED == 1) X = red_near60(X, nq);
                        else if (RED == 2) X = red_csub(X, eightq);
                        else if (RED == 3) X = red_mask(X, delta);
                    }
                    const u64 T = V == 0 ? shoup4_asm(x[e + stp], w.w, w.ws, nq) : V == 1 ? shoup4_c(x[e + stp], w.w, w.ws, nq)
                                                                                          : shoup4_near(x[e + stp], w.w, w.ws, nq);
                    x[e] = X + T; x[e + stp] = X + fourq - T;
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 16; e++) s[(e << 8) | tid] = x[e];
        __syncthreads();
    }
    for (int i = 0; i < 16; i++) g[(blockIdx.x & 1023) * 4096 + tid + 256 * i] = s[tid + 256 * i];
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

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    u64* buf; cudaMalloc(&buf, 1024 * 4096 * 8); cudaMemset(buf, 1, 1024 * 4096 * 8);
    Tw h[16]; for (int i = 0; i < 16; i++) { h[i].w = 0x123456789abcdefULL + i; h[i].ws = 0x23456789abcdef01ULL * (i + 1); }
    Tw* dt; cudaMalloc(&dt, sizeof(h)); cudaMemcpy(dt, h, sizeof(h), cudaMemcpyHostToDevice);
    const u64 q = 0xffffffffffc0001ULL;
    const int rounds = 256;
#define RUN(V, RED, MINB)                                                                                                   \
    {                                                                                                                       \
        const int blocks = sms * MINB * 4;                                                                                  \
        const float ms = time_ms([&] { k<V, RED, MINB><<<blocks, 256>>>(buf, dt, q, rounds); });                            \
        const double bf = (double)blocks * 256 * 32 * rounds;                                                               \
        printf("{\"name\": \"round16_v%d_red%d_occ%d\", \"ms\": %.4f, \"Gbfly_per_s\": %.1f, \"cycles_per_warp_bfly_smsp\": %.2f}\n", V, RED, \
               MINB, ms, bf / ms / 1e6, (double)sms * 4 * (p.clockRate * 1e3) * (ms * 1e-3) / (bf / 32));                   \
    }
    RUN(0, 0, 2) RUN(0, 0, 3) RUN(0, 0, 4)
    RUN(1, 0, 2) RUN(1, 0, 3) RUN(1, 0, 4)
    RUN(2, 0, 2) RUN(2, 0, 3) RUN(2, 0, 4)
    RUN(0, 1, 3) RUN(0, 2, 3) RUN(0, 3, 3)
    RUN(1, 1, 3) RUN(1, 2, 3) RUN(1, 3, 3)
    RUN(2, 1, 3) RUN(2, 2, 3) RUN(2, 3, 3)
    printf("{\"cuda_error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
