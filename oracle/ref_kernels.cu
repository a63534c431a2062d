// Harness around the REFERENCE's own arithmetic kernels (test infrastructure; never linked into the product).
// Compiled only where /root/reference is present (oracle/Makefile target _ref), from the reference sources where they lie:
// this file #includes REF_BIGINT_CU = /root/reference/src/bigint.cu, i.e. fhe::batch_mod_add_kernel / batch_mod_sub_kernel /
// batch_mod_mul_kernel (src/bigint.cu:171-215) with fhe::add_mod / sub_mod / mul_mod_montgomery (include/bigint.cuh:27-140) and the
// host fhe::compute_montgomery_inverse (src/bigint.cu:23-39).  These are the only device functions of the reference whose
// definitions are complete; everything else there is a placeholder.  The output, oracle/_ref/libref_kernels.so,
// is git-ignored and travels to the GPU box with the snapshot.  Nothing below is reference code: it widens uint64 operands to
// the reference's uint256_t, launches the reference's kernels as its own test does (tests/test_fhe.cu:43,51: 256 threads per
// block), and narrows the results.
#include <cstdint>
#include <vector>
#include REF_BIGINT_CU

namespace {
int run(int op, const uint64_t* h_a, const uint64_t* h_b, uint64_t q, uint64_t* h_out, uint32_t count, uint64_t* h_hi_or) {
    using fhe::uint256_t;
    std::vector<uint256_t> a(count), b(count), r(count);
    for (uint32_t i = 0; i < count; i++) { a[i] = uint256_t(h_a[i]); b[i] = uint256_t(h_b[i]); }
    uint256_t *d_a = nullptr, *d_b = nullptr, *d_r = nullptr;
    const size_t bytes = (size_t)count * sizeof(uint256_t);
    if (cudaMalloc(&d_a, bytes) != cudaSuccess || cudaMalloc(&d_b, bytes) != cudaSuccess || cudaMalloc(&d_r, bytes) != cudaSuccess) return -1;
    cudaMemcpy(d_a, a.data(), bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(d_b, b.data(), bytes, cudaMemcpyHostToDevice);
    const uint256_t mod(q);
    const uint32_t threads = 256, blocks = (count + threads - 1) / threads;
    if (op == 0) fhe::batch_mod_add_kernel<<<blocks, threads>>>(d_r, d_a, d_b, mod, count);
    else if (op == 1) fhe::batch_mod_sub_kernel<<<blocks, threads>>>(d_r, d_a, d_b, mod, count);
    else fhe::batch_mod_mul_kernel<<<blocks, threads>>>(d_r, d_a, d_b, mod, fhe::compute_montgomery_inverse(mod), count);
    const cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(r.data(), d_r, bytes, cudaMemcpyDeviceToHost);
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_r);
    if (e != cudaSuccess) return -2;
    uint64_t hi = 0;
    for (uint32_t i = 0; i < count; i++) { h_out[i] = r[i].limbs[0]; hi |= r[i].limbs[1] | r[i].limbs[2] | r[i].limbs[3]; }
    if (h_hi_or) *h_hi_or = hi;                   // non-zero if any result did not fit 64 bits
    return 0;
}
}  // namespace

extern "C" int ref_batch_mod_add(const uint64_t* a, const uint64_t* b, uint64_t q, uint64_t* out, uint32_t n, uint64_t* hi) { return run(0, a, b, q, out, n, hi); }
extern "C" int ref_batch_mod_sub(const uint64_t* a, const uint64_t* b, uint64_t q, uint64_t* out, uint32_t n, uint64_t* hi) { return run(1, a, b, q, out, n, hi); }
// reference Montgomery product: a * b * 2^-256 mod q (the reference never converts into the Montgomery domain)
extern "C" int ref_batch_mod_mul_montgomery(const uint64_t* a, const uint64_t* b, uint64_t q, uint64_t* out, uint32_t n, uint64_t* hi) { return run(2, a, b, q, out, n, hi); }

// Timing of the same kernels on device-resident operands (CUDA events, `reps` launches after one warm-up): the "reference CUDA
// build on the same box" data point of SURVEY 8d for what can legally launch.  Operands are pseudo-random 60-bit values below q.
extern "C" int ref_time_batch_kernel(int op, uint64_t q, uint32_t count, int reps, float* ms_per_launch) {
    using fhe::uint256_t;
    std::vector<uint256_t> a(count), b(count);
    uint64_t x = 0x9E3779B97F4A7C15ull;
    for (uint32_t i = 0; i < count; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17; a[i] = uint256_t(x % q);
        x ^= x << 13; x ^= x >> 7; x ^= x << 17; b[i] = uint256_t(x % q);
    }
    uint256_t *d_a = nullptr, *d_b = nullptr, *d_r = nullptr;
    const size_t bytes = (size_t)count * sizeof(uint256_t);
    if (cudaMalloc(&d_a, bytes) != cudaSuccess || cudaMalloc(&d_b, bytes) != cudaSuccess || cudaMalloc(&d_r, bytes) != cudaSuccess) return -1;
    cudaMemcpy(d_a, a.data(), bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(d_b, b.data(), bytes, cudaMemcpyHostToDevice);
    const uint256_t mod(q), inv = fhe::compute_montgomery_inverse(mod);
    const uint32_t threads = 256, blocks = (count + threads - 1) / threads;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = -1; r < reps; r++) {
        if (r == 0) cudaEventRecord(e0);
        if (op == 0) fhe::batch_mod_add_kernel<<<blocks, threads>>>(d_r, d_a, d_b, mod, count);
        else if (op == 1) fhe::batch_mod_sub_kernel<<<blocks, threads>>>(d_r, d_a, d_b, mod, count);
        else fhe::batch_mod_mul_kernel<<<blocks, threads>>>(d_r, d_a, d_b, mod, inv, count);
    }
    cudaEventRecord(e1);
    const cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_a); cudaFree(d_b); cudaFree(d_r);
    if (e != cudaSuccess) return -2;
    *ms_per_launch = ms / reps;
    return 0;
}
