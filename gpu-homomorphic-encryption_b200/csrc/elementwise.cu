// Fused element-wise polynomial kernels over [batch][limb_count][N] (HBM-bound: 16..32 bytes per coefficient).
// Replaces poly_add/sub/mul_scalar kernels (/root/reference/src/polynomial.cu:70-111), ntt_pointwise_mul_kernel
// (/root/reference/kernels/ntt_kernels.cu:124-137), rns_add/rns_mul kernels (/root/reference/src/rns.cu:143-180),
// bit_reverse_kernel (kernels/ntt_kernels.cu:140-161) and the uint256_t edge (include/bigint.cuh:9-24).
// One thread moves 16 bytes per operand per iteration (LDG.128/STG.128); grids are sized to a multiple of the SM count.
#include "common.cuh"
#include "host_math.hpp"

namespace fhe_b200 {

// per-limb scalars travel as a kernel parameter (no device scratch, no cross-stream hazards)
struct ScalarPack { u64 v[64]; };

template <int OP>
__global__ void __launch_bounds__(256) ew_kernel(ulonglong2* __restrict__ out, const ulonglong2* __restrict__ a,
                                                 const ulonglong2* __restrict__ b, const ulonglong2* __restrict__ c,
                                                 const LimbParams* __restrict__ params, const ScalarPack scalars,
                                                 uint32_t logn, uint32_t limb_begin, uint32_t limb_count, size_t nvec) {
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        const uint32_t l = (uint32_t)(((2 * v) >> logn) % limb_count);
        const LimbParams P = params[limb_begin + l];
        const ulonglong2 x = a[v];
        ulonglong2 r;
        if (OP == EW_ADD) { const ulonglong2 y = b[v]; r.x = add_mod(x.x, y.x, P.q); r.y = add_mod(x.y, y.y, P.q); }
        if (OP == EW_SUB) { const ulonglong2 y = b[v]; r.x = sub_mod(x.x, y.x, P.q); r.y = sub_mod(x.y, y.y, P.q); }
        if (OP == EW_MUL) { const ulonglong2 y = b[v]; r.x = mul_mod(x.x, y.x, P); r.y = mul_mod(x.y, y.y, P); }
        if (OP == EW_MAC) {   // out = c + a*b
            const ulonglong2 y = b[v]; const ulonglong2 z = c[v];
            r.x = add_mod(z.x, mul_mod(x.x, y.x, P), P.q); r.y = add_mod(z.y, mul_mod(x.y, y.y, P), P.q);
        }
        if (OP == EW_MUL_SCALAR) { const u64 s = scalars.v[l]; r.x = mul_mod(x.x, s, P); r.y = mul_mod(x.y, s, P); }
        if (OP == EW_ADD_SCALAR) { const u64 s = scalars.v[l]; r.x = add_mod(x.x, s, P.q); r.y = add_mod(x.y, s, P.q); }
        if (OP == EW_NEG) { r.x = neg_mod(x.x, P.q); r.y = neg_mod(x.y, P.q); }
        out[v] = r;
    }
}

static inline uint32_t ew_grid(const fhe_b200_plan* plan, size_t nvec) {
    const size_t want = (nvec + 255) / 256;
    const size_t cap = (size_t)plan->sm_count * 16;     // 16 resident CTAs of 256 threads per SM is plenty for a stream kernel
    return (uint32_t)(want < cap ? want : cap);
}

int launch_elementwise(fhe_b200_plan* plan, EwOp op, uint64_t* d_out, const uint64_t* d_a, const uint64_t* d_b,
                       const uint64_t* d_c, uint32_t batch, uint32_t limb_begin, uint32_t limb_count, cudaStream_t st) {
    FHE_TRY(check_range(plan, batch, limb_begin, limb_count));
    const size_t nvec = (size_t)batch * limb_count * plan->n / 2;
    if (nvec == 0) return 0;
    DeviceGuard dev_guard(plan->device);
    const uint32_t grid = ew_grid(plan, nvec);
    ScalarPack sc;
    if (op == EW_MUL_SCALAR || op == EW_ADD_SCALAR) {
        // d_b carries a HOST array of limb_count residues for the scalar ops
        FHE_REQUIRE(limb_count <= 64, "scalar ops take at most 64 limbs per call");
        for (uint32_t l = 0; l < limb_count; l++) sc.v[l] = d_b[l] % plan->moduli[limb_begin + l];
        d_b = nullptr;
    }
    auto o = reinterpret_cast<ulonglong2*>(d_out);
    auto a = reinterpret_cast<const ulonglong2*>(d_a);
    auto b = reinterpret_cast<const ulonglong2*>(d_b);
    auto c = reinterpret_cast<const ulonglong2*>(d_c);
#define EW_LAUNCH(OPV) ew_kernel<OPV><<<grid, 256, 0, st>>>(o, a, b, c, plan->d_params, sc, plan->logn, limb_begin, limb_count, nvec)
    switch (op) {
        case EW_ADD: EW_LAUNCH(EW_ADD); break;
        case EW_SUB: EW_LAUNCH(EW_SUB); break;
        case EW_MUL: EW_LAUNCH(EW_MUL); break;
        case EW_MAC: EW_LAUNCH(EW_MAC); break;
        case EW_MUL_SCALAR: EW_LAUNCH(EW_MUL_SCALAR); break;
        case EW_ADD_SCALAR: EW_LAUNCH(EW_ADD_SCALAR); break;
        case EW_NEG: EW_LAUNCH(EW_NEG); break;
    }
#undef EW_LAUNCH
    FHE_LAUNCH_CHECK();
    return 0;
}

// ---- bit-reversal permutation: out[bitrev(k)] = in[k], per polynomial-limb of N elements -----------------------
__global__ void __launch_bounds__(256) bitrev_kernel(u64* __restrict__ out, const u64* __restrict__ in, uint32_t logn,
                                                     size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t k = (uint32_t)(i & ((1u << logn) - 1));
        const size_t base = i - k;
        out[base + (__brev(k) >> (32 - logn))] = in[i];
    }
}

// ---- fhe::uint256_t edge ---------------------------------------------------------------------------------------
// uint256_t = 4 little-endian u64 words; two 16-byte halves per value
__global__ void __launch_bounds__(256) unpack_u256_kernel(u64* __restrict__ out, const ulonglong2* __restrict__ in, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        out[i] = in[2 * i].x;
}
__global__ void __launch_bounds__(256) pack_u256_kernel(ulonglong2* __restrict__ out, const u64* __restrict__ in, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        out[2 * i] = make_ulonglong2(in[i], 0ull);
        out[2 * i + 1] = make_ulonglong2(0ull, 0ull);
    }
}
// residue of a 256-bit value modulo q: Horner over the four words with 2^64 mod q
__global__ void __launch_bounds__(256) to_rns_u256_kernel(u64* __restrict__ out, const ulonglong2* __restrict__ in,
                                                          const LimbParams* __restrict__ params, uint32_t limb_begin,
                                                          uint32_t limb_count, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const ulonglong2 lo = in[2 * i], hi = in[2 * i + 1];
        for (uint32_t l = 0; l < limb_count; l++) {
            const LimbParams P = params[limb_begin + l];
            u64 r = barrett128(hi.y, hi.x, P.q, P.mu_hi, P.mu_lo);      // (w3:w2) mod q
            r = barrett128(r, lo.y, P.q, P.mu_hi, P.mu_lo);             // (r:w1) mod q
            r = barrett128(r, lo.x, P.q, P.mu_hi, P.mu_lo);             // (r:w0) mod q
            out[(size_t)l * count + i] = r;
        }
    }
}

// Garner mixed-radix CRT into 256 bits: x = v0 + v1 q0 + v2 q0 q1 + v3 q0 q1 q2, v_i < q_i.
struct GarnerConsts { u64 q[4]; u64 inv[4][4]; };     // inv[i][j] = (q_j)^-1 mod q_i for j < i
__global__ void __launch_bounds__(256) from_rns_u256_kernel(ulonglong2* __restrict__ out, const u64* __restrict__ in,
                                                            const LimbParams* __restrict__ params, uint32_t limb_begin,
                                                            uint32_t k, size_t count, const GarnerConsts gc) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        u64 v[4] = {0, 0, 0, 0};
        for (uint32_t a = 0; a < k; a++) {
            const LimbParams P = params[limb_begin + a];
            u64 t = in[(size_t)a * count + i];
            // t = (...((x_a - v0) q0^-1 - v1) q1^-1 ...) mod q_a
            for (uint32_t b = 0; b < a; b++) {
                const u64 vb = barrett128(0, v[b], P.q, P.mu_hi, P.mu_lo);
                t = mul_mod(sub_mod(t, vb, P.q), gc.inv[a][b], P);
            }
            v[a] = t;
        }
        // Horner: ((v3 q2 + v2) q1 + v1) q0 + v0 over four 64-bit words
        u64 w[4] = {0, 0, 0, 0};
        for (int a = (int)k - 1; a >= 0; a--) {
            if (a < (int)k - 1) {           // w *= q[a]
                u64 carry = 0;
                for (int j = 0; j < 4; j++) {
                    u64 hi, lo; mul128(w[j], gc.q[a], hi, lo);
                    lo += carry; hi += (lo < carry);
                    w[j] = lo; carry = hi;
                }
            }
            u64 c = v[a];                    // w += v[a]
            for (int j = 0; j < 4 && c; j++) { w[j] += c; c = (w[j] < c) ? 1 : 0; }
        }
        out[2 * i] = make_ulonglong2(w[0], w[1]);
        out[2 * i + 1] = make_ulonglong2(w[2], w[3]);
    }
}

static inline uint32_t flat_grid(size_t n) { const size_t w = (n + 255) / 256; return (uint32_t)(w < 148 * 16 ? (w ? w : 1) : 148 * 16); }

}  // namespace fhe_b200

using namespace fhe_b200;

extern "C" int fhe_b200_bitrev_permute(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t n_polys, void* stream) {
    FHE_REQUIRE(plan && d_out && d_in && d_out != d_in, "bitrev_permute: null or aliased buffers");
    const size_t total = (size_t)n_polys * plan->n;
    if (!total) return 0;
    DeviceGuard dev_guard(plan->device);
    bitrev_kernel<<<flat_grid(total), 256, 0, (cudaStream_t)stream>>>(d_out, d_in, plan->logn, total);
    FHE_LAUNCH_CHECK();
    return 0;
}
extern "C" int fhe_b200_unpack_u256(uint64_t* d_out, const void* d_u256, size_t count, void* stream) {
    FHE_REQUIRE(d_out && d_u256, "unpack_u256: null buffer");
    if (!count) return 0;
    unpack_u256_kernel<<<flat_grid(count), 256, 0, (cudaStream_t)stream>>>(d_out, (const ulonglong2*)d_u256, count);
    FHE_LAUNCH_CHECK();
    return 0;
}
extern "C" int fhe_b200_pack_u256(void* d_u256, const uint64_t* d_in, size_t count, void* stream) {
    FHE_REQUIRE(d_in && d_u256, "pack_u256: null buffer");
    if (!count) return 0;
    pack_u256_kernel<<<flat_grid(count), 256, 0, (cudaStream_t)stream>>>((ulonglong2*)d_u256, d_in, count);
    FHE_LAUNCH_CHECK();
    return 0;
}
extern "C" int fhe_b200_from_rns_u256(fhe_b200_plan* plan, void* d_u256, const uint64_t* d_in, size_t count, uint32_t limb_begin,
                                      uint32_t limb_count, void* stream) {
    FHE_REQUIRE(plan && d_u256 && d_in, "from_rns_u256: null argument");
    FHE_TRY(check_range(plan, 1, limb_begin, limb_count));
    FHE_REQUIRE(limb_count >= 1 && limb_count <= 4, "from_rns_u256: 1..4 limbs (Q must fit 256 bits)");
    double bits = 0;
    for (uint32_t a = 0; a < limb_count; a++) { double b = 0; for (uint64_t t = plan->h_params[limb_begin + a].q; t; t >>= 1) b++; bits += b; }
    FHE_REQUIRE(bits <= 256, "from_rns_u256: the product of the moduli exceeds 256 bits");
    if (!count) return 0;
    DeviceGuard dev_guard(plan->device);
    GarnerConsts gc;
    for (uint32_t a = 0; a < 4; a++) {
        gc.q[a] = a < limb_count ? plan->moduli[limb_begin + a] : 1;
        for (uint32_t b = 0; b < 4; b++) {
            gc.inv[a][b] = 0;
            if (a < limb_count && b < a) gc.inv[a][b] = host::invmod(plan->moduli[limb_begin + b] % gc.q[a], gc.q[a]);
        }
    }
    from_rns_u256_kernel<<<flat_grid(count), 256, 0, (cudaStream_t)stream>>>((ulonglong2*)d_u256, d_in, plan->d_params, limb_begin,
                                                                            limb_count, count, gc);
    FHE_LAUNCH_CHECK();
    return 0;
}
extern "C" int fhe_b200_to_rns_u256(fhe_b200_plan* plan, uint64_t* d_out, const void* d_u256, size_t count,
                                    uint32_t limb_begin, uint32_t limb_count, void* stream) {
    FHE_REQUIRE(plan && d_out && d_u256, "to_rns_u256: null argument");
    FHE_TRY(check_range(plan, 1, limb_begin, limb_count));
    if (!count) return 0;
    DeviceGuard dev_guard(plan->device);
    to_rns_u256_kernel<<<flat_grid(count), 256, 0, (cudaStream_t)stream>>>(d_out, (const ulonglong2*)d_u256, plan->d_params,
                                                                          limb_begin, limb_count, count);
    FHE_LAUNCH_CHECK();
    return 0;
}
