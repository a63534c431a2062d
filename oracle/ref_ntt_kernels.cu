// Harness around the REFERENCE's NTT / polynomial kernels for TIMING on the same box (test infrastructure; never linked into the
// product).  Compiled only where /root/reference is present (oracle/Makefile target _ref) from the reference sources where they
// lie: this file #includes REF_NTT_KERNELS_CU = /root/reference/kernels/ntt_kernels.cu (ntt_forward_optimized_kernel :7-62,
// ntt_pointwise_mul_kernel :124-137, bit_reverse_kernel :140-161) and REF_POLY_CU = /root/reference/src/polynomial.cu
// (poly_add_kernel :70-82).  The launches below repeat the reference's own launch configurations:
//   NTTEngine::forward   src/ntt.cu:30-40   bit_reverse_kernel<<<ceil(n/256),256>>> then ntt_forward_optimized_kernel<<<1, n, n*32 B>>>
//                        -- legal only for n <= 1024 (one block of n threads), which is why N = 1024 is the only size timed;
//   NTTEngine::multiply  src/ntt.cu:65-68   ntt_pointwise_mul_kernel<<<ceil(n/256),256>>>
//   PolynomialOps::add   src/polynomial.cu:36-42   poly_add_kernel<<<ceil(n/256),256>>>
// The reference's twiddle tables are placeholders (1, 2, 3, ...; src/ntt.cu:86-97), so the transform's OUTPUT is meaningless and is
// not compared with anything: only its duration is reported.  src/ntt.cu itself does not compile (F2).
// The member functions of src/polynomial.cu that call into files which do not compile are never called here; the two symbols they
// reference are given empty bodies below so that the shared object loads.
#include <cstdint>
#include <vector>
#include REF_NTT_KERNELS_CU
#include REF_POLY_CU

namespace fhe {      // never called: only to satisfy the loader (src/ntt.cu does not compile)
void NTTEngine::multiply(uint256_t*, const uint256_t*, const uint256_t*) {}
MontgomeryParams compute_montgomery_params(const uint256_t&) { return MontgomeryParams(); }
}

namespace {
using fhe::uint256_t;
uint256_t* device_values(uint32_t count, uint64_t q, uint64_t seed) {
    std::vector<uint256_t> a(count);
    uint64_t x = seed;
    for (uint32_t i = 0; i < count; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; a[i] = uint256_t(x % q); }
    uint256_t* d = nullptr;
    if (cudaMalloc(&d, (size_t)count * sizeof(uint256_t)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, a.data(), (size_t)count * sizeof(uint256_t), cudaMemcpyHostToDevice);
    return d;
}
uint256_t mont_inv_of(uint64_t q) {      // -q^-1 mod 2^64 in the low word (what compute_montgomery_inverse produces, src/bigint.cu:23-39)
    uint64_t inv = q;
    for (int i = 0; i < 6; i++) inv *= 2 - q * inv;
    return uint256_t(0 - inv);
}
}  // namespace

// what: 0 = ntt_pointwise_mul_kernel, 1 = poly_add_kernel over `count` coefficients (one launch = one polynomial, as the reference does);
//       2 = NTTEngine::forward's two launches at n = count (count <= 1024).  ms per call, `reps` calls after one warm-up.
extern "C" int ref_time_poly_kernel(int what, uint64_t q, uint32_t count, int reps, float* ms_per_call) {
    if (what == 2 && count > 1024) return -3;                    // illegal launch in the reference (src/ntt.cu:36-39)
    uint256_t *d_a = device_values(count, q, 0x9E3779B97F4A7C15ull), *d_b = device_values(count, q, 0xD1B54A32D192ED03ull), *d_r = nullptr;
    if (!d_a || !d_b || cudaMalloc(&d_r, (size_t)count * sizeof(uint256_t)) != cudaSuccess) return -1;
    const uint256_t mod(q), inv = mont_inv_of(q);
    const uint32_t blocks = (count + 255) / 256;
    if (what == 2) cudaFuncSetAttribute(fhe::ntt_forward_optimized_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(count * sizeof(uint256_t)));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = -1; r < reps; r++) {
        if (r == 0) cudaEventRecord(e0);
        if (what == 0) fhe::ntt_pointwise_mul_kernel<<<blocks, 256>>>(d_r, d_a, d_b, mod, inv, count);
        else if (what == 1) fhe::poly_add_kernel<<<blocks, 256>>>(d_r, d_a, d_b, mod, count);
        else {
            fhe::bit_reverse_kernel<<<blocks, 256>>>(d_a, count);
            fhe::ntt_forward_optimized_kernel<<<1, count, count * sizeof(uint256_t)>>>(d_a, d_b /* twiddles: any table */, mod, inv, count);
        }
    }
    cudaEventRecord(e1);
    const cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_a); cudaFree(d_b); cudaFree(d_r);
    if (e != cudaSuccess) return -2;
    *ms_per_call = ms / reps;
    return 0;
}
