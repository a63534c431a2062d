// Does the legacy warp-level int8 MMA (mma.sync.m16n8k32.u8.u8.s32 -> IMMA) run at a useful rate on B200, and is the
// fragment layout what we think it is?  Prints a JSON line: layout check against a CPU product + sustained TOPS.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_u8(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// A: 16x32 row-major bytes, B: 32x8 (k-major per column: B[k][n]), C: 16x8 ints
__global__ void k_check(const uint8_t* A, const uint8_t* B, int* C) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    uint32_t a[4], b[2]; int c[4] = {0, 0, 0, 0};
    auto ldA = [&](int row, int col) { uint32_t v = 0; for (int i = 0; i < 4; i++) v |= (uint32_t)A[row * 32 + col + i] << (8 * i); return v; };
    auto ldB = [&](int k, int n) { uint32_t v = 0; for (int i = 0; i < 4; i++) v |= (uint32_t)B[(k + i) * 8 + n] << (8 * i); return v; };
    a[0] = ldA(g, t * 4); a[1] = ldA(g + 8, t * 4); a[2] = ldA(g, 16 + t * 4); a[3] = ldA(g + 8, 16 + t * 4);
    b[0] = ldB(t * 4, g); b[1] = ldB(16 + t * 4, g);
    mma_u8(c, a, b);
    C[g * 8 + t * 2] = c[0]; C[g * 8 + t * 2 + 1] = c[1]; C[(g + 8) * 8 + t * 2] = c[2]; C[(g + 8) * 8 + t * 2 + 1] = c[3];
}

__global__ void k_rate(int* out, int iters) {
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x * 5u, 11u};
    int c[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) mma_u8(c[i], a, b);
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j];
    if (s == 0x12345678) out[0] = s;
}

int main() {
    uint8_t hA[16 * 32], hB[32 * 8]; int hC[16 * 8], ref[16 * 8];
    srand(1);
    for (auto& v : hA) v = rand() & 255;
    for (auto& v : hB) v = rand() & 255;
    for (int m = 0; m < 16; m++) for (int n = 0; n < 8; n++) { int s = 0; for (int k = 0; k < 32; k++) s += (int)hA[m * 32 + k] * (int)hB[k * 8 + n]; ref[m * 8 + n] = s; }
    uint8_t *dA, *dB; int* dC;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dC, sizeof(hC));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    k_check<<<1, 32>>>(dA, dB, dC);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hC, dC, sizeof(hC), cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < 128; i++) bad += hC[i] != ref[i];
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_rate<<<blocks, threads>>>(dC, iters); cudaDeviceSynchronize();
    cudaEventRecord(a); k_rate<<<blocks, threads>>>(dC, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double mmas = (double)blocks * (threads / 32) * iters * 8;
    printf("{\"layout_mismatches\": %d, \"cuda_error\": \"%s\", \"imma_m16n8k32_TOPS\": %.1f, \"mma_per_clk_per_sm\": %.3f}\n", bad, cudaGetErrorString(e),
           mmas * 16 * 8 * 32 * 2 / ms / 1e9, mmas / (ms * 1e-3) / (p.clockRate * 1e3) / p.multiProcessorCount);
    return 0;
}
