// __global__ wrappers and launchers of the batched negacyclic NTT (bodies in ntt_core.cuh).
//
// Work items over a batch laid out [batch][limb_count][N]:
//   tile item : (limb, tile, polynomial group).  A CTA of 2^(LB-4) threads stages that tile's twiddles in shared memory once
//               (two cp.async.bulk copies completing on an mbarrier) and runs the tile pass for every polynomial of the group,
//               prefetching the next polynomial's tile while it finishes the current one.
//   row item  : (limb, polynomial, column block): V adjacent columns x 2^K1 rows per thread, registers only.
//
// N <= 4096: tile items only (one kernel).  N > 4096 (where the balanced passes of ntt_bal.cu do not apply): row kernel then tile
// kernel (per chunk of plan->chunk_bytes).  Both passes are bound by the IMAD pipe at the same per-stage rate (measured).  (A
// persistent single-kernel variant with release/acquire counters between row and tile items was measured 5% slower at config 3 in
// round 1 and has been removed.)
#include "common.cuh"
#include "ntt_core.cuh"
#include "tma.cuh"

namespace fhe_b200 {

struct NttArgs {
    uint64_t* out;
    const uint64_t* in;
    const Twiddle* tw;           // [limbs][n] table of the direction in use (row pass)
    const Twiddle* p12;          // [limbs][tiles][256] staged blocks of the tile pass
    const Twiddle* p3;           // [limbs][tiles][p3n]
    const LimbParams* params;    // [limbs]
    uint32_t n;                  // ring degree
    uint32_t limb_count;         // limbs per polynomial in the buffer
    uint32_t limb_begin;         // first plan limb the buffer's limb 0 maps to
    uint32_t l0, nl;             // chunk: buffer limbs [l0, l0+nl)
    uint32_t b0, nb;             // chunk: polynomials [b0, b0+nb)
    uint32_t groups;             // two-pass tile kernel: CTAs per (limb, tile); CTA g handles polynomials b0+g, b0+g+groups, ...
    uint32_t tiles, p3n;
};

// (limb, tile/column-block, polynomial or group) from a linear CTA index, innermost fastest
__device__ __forceinline__ void decode(uint32_t i, uint32_t inner, uint32_t units, uint32_t l0, uint32_t& limb, uint32_t& unit,
                                       uint32_t& in_idx) {
    in_idx = i % inner;
    const uint32_t r = i / inner;
    unit = r % units;
    limb = l0 + r / units;
}

// 16-byte asynchronous global -> shared copies (LDGSTS, L2 only) of one tile into the swizzled layout
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
template <int LB>
__device__ __forceinline__ void tile_copy_in_async(uint32_t tid, const u64* g, u64* s) {
    constexpr int NT = 1 << (LB - 4);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t c = tid + i * NT;                            // logical 16-byte chunk
        const uint32_t p = ((c >> 3) << 3) | ((c & 7) ^ ((c >> 3) & 7));
        cp_async_16(s + 2 * p, g + 2 * c);
    }
}

template <int LB>
struct TileSmem {
    static constexpr int NB = 1 << LB;
    static constexpr size_t data_bytes = (size_t)NB * 8;
    static constexpr size_t p12_bytes = 256 * sizeof(Twiddle);
    static constexpr size_t p3_bytes = (size_t)p3_entries(LB) * sizeof(Twiddle);
    static constexpr size_t total = data_bytes + p12_bytes + p3_bytes + 32;
    __device__ static u64* data(unsigned char* raw) { return reinterpret_cast<u64*>(raw); }
    __device__ static Twiddle* p12(unsigned char* raw) { return reinterpret_cast<Twiddle*>(raw + data_bytes); }
    __device__ static Twiddle* p3(unsigned char* raw) { return reinterpret_cast<Twiddle*>(raw + data_bytes + p12_bytes); }
    __device__ static uint64_t* bar(unsigned char* raw) { return reinterpret_cast<uint64_t*>(raw + data_bytes + p12_bytes + p3_bytes); }
    __device__ static uint32_t* scratch(unsigned char* raw) { return reinterpret_cast<uint32_t*>(raw + data_bytes + p12_bytes + p3_bytes + 16); }
};

template <int LB>
__device__ __forceinline__ void stage_twiddles(const NttArgs& a, unsigned char* raw, uint32_t pl, uint32_t tile) {
    using SM = TileSmem<LB>;
    if (threadIdx.x == 0) {
        uint64_t* bar = SM::bar(raw);
        mbar_arrive_expect_tx(bar, (uint32_t)(SM::p12_bytes + SM::p3_bytes));
        bulk_copy_g2s(SM::p12(raw), a.p12 + ((size_t)pl * a.tiles + tile) * 256, (uint32_t)SM::p12_bytes, bar);
        bulk_copy_g2s(SM::p3(raw), a.p3 + ((size_t)pl * a.tiles + tile) * a.p3n, (uint32_t)SM::p3_bytes, bar);
    }
}

// ---- tile item, forward: polynomials poly, poly+step, ... < poly_end of (limb, tile).  The mbarrier must be initialised;
// `parity` is the barrier phase this item completes (the caller flips it afterwards).
template <int LB, int K1, int HB, bool NEAR>
__device__ __forceinline__ void tile_fwd_item(const NttArgs& a, unsigned char* raw, uint32_t limb, uint32_t tile, uint32_t poly,
                                              uint32_t poly_end, uint32_t step, uint32_t parity) {
    using T = TileFwd<LB, HB, NEAR>;
    using SM = TileSmem<LB>;
    u64* s = SM::data(raw);
    const Twiddle* s12 = SM::p12(raw);
    const Twiddle* s3 = SM::p3(raw);
    const uint32_t pl = a.limb_begin + limb;
    const uint32_t tid = threadIdx.x;
    stage_twiddles<LB>(a, raw, pl, tile);
    const LimbParams P = a.params[pl];
    constexpr int B0 = fwd_bound_after(1, K1, HB, NEAR);
    const u64* base_in = (K1 > 0 ? a.out : a.in);             // K1 > 0: the row pass already moved the data to `out`
    const size_t limb_off = (size_t)limb * a.n + (size_t)tile * T::NB;
    const size_t poly_stride = (size_t)a.limb_count * a.n;
    u64 x[16];
    if (poly < poly_end) T::phase1_load(tid, base_in + poly * poly_stride + limb_off, x);
    mbar_wait(SM::bar(raw), parity);                          // staged twiddles have landed
    for (; poly < poly_end; poly += step) {
        const size_t off = poly * poly_stride + limb_off;
        T::template phase1_compute<B0>(tid, x, s, s12, P);
        __syncthreads();
        T::template phase2<B0>(tid, s, s12, P);
        __syncthreads();
        T::template phase3<B0>(tid, s, s3, P);
        __syncthreads();
        // software pipeline: the next polynomial's global loads fly while this one is copied out
        const uint32_t next = poly + step;
        if (next < poly_end) T::phase1_load(tid, base_in + next * poly_stride + limb_off, x);
        T::phase4(tid, a.out + off, s);
        __syncthreads();
    }
}

template <int LB, int K1, int HB, bool NEAR>
__device__ __forceinline__ void tile_inv_item(const NttArgs& a, unsigned char* raw, uint32_t limb, uint32_t tile, uint32_t poly,
                                              uint32_t poly_end, uint32_t step, uint32_t parity) {
    using T = TileInv<LB, HB, NEAR>;
    using SM = TileSmem<LB>;
    u64* s = SM::data(raw);
    const Twiddle* s12 = SM::p12(raw);
    const Twiddle* s3 = SM::p3(raw);
    const uint32_t pl = a.limb_begin + limb;
    const uint32_t tid = threadIdx.x;
    stage_twiddles<LB>(a, raw, pl, tile);
    const LimbParams P = a.params[pl];
    const size_t limb_off = (size_t)limb * a.n + (size_t)tile * T::NB;
    const size_t poly_stride = (size_t)a.limb_count * a.n;
    if (poly < poly_end) tile_copy_in_async<LB>(tid, a.in + poly * poly_stride + limb_off, s);
    mbar_wait(SM::bar(raw), parity);
    for (; poly < poly_end; poly += step) {
        const size_t off = poly * poly_stride + limb_off;
        cp_async_wait_all();
        __syncthreads();
        T::phase2(tid, s, s3, P);
        __syncthreads();
        T::phase3(tid, s, s12, P);
        __syncthreads();
        u64 x[16];
        T::phase4_load(tid, s, x);
        __syncthreads();                                       // every thread has its inputs: the tile buffer is free
        const uint32_t next = poly + step;
        if (next < poly_end) tile_copy_in_async<LB>(tid, a.in + next * poly_stride + limb_off, s);   // overlaps the last 4 stages
        T::template phase4_compute<K1 == 0>(tid, x, a.out + off, s12, P);
    }
    __syncthreads();                                           // twiddle blocks may be restaged after this
}

constexpr int kRowThreads = 256;

template <int LB, int K1, int HB, bool NEAR, int V>
__device__ __forceinline__ void row_fwd_item(const NttArgs& a, uint32_t limb, uint32_t poly, uint32_t cb) {
    const size_t off = ((size_t)poly * a.limb_count + limb) * a.n;
    const uint32_t pl = a.limb_begin + limb;
    const uint32_t col = (cb * kRowThreads + threadIdx.x) * V;
    RowPass<K1, V, LB, HB, NEAR>::forward(a.out + off, a.in + off, col, a.tw + (size_t)pl * a.n, a.params[pl]);
}
template <int LB, int K1, int HB, bool NEAR, int V>
__device__ __forceinline__ void row_inv_item(const NttArgs& a, uint32_t limb, uint32_t poly, uint32_t cb) {
    const size_t off = ((size_t)poly * a.limb_count + limb) * a.n;
    const uint32_t pl = a.limb_begin + limb;
    const uint32_t col = (cb * kRowThreads + threadIdx.x) * V;
    constexpr int B0 = TileInv<LB, HB, NEAR>::out_bound();
    // in place on `out`: the tile pass wrote there
    RowPass<K1, V, LB, HB, NEAR>::template inverse<B0>(a.out + off, col, a.tw + (size_t)pl * a.n, a.params[pl]);
}

// ================================================ two-pass kernels =====================================================
template <int LB, int K1, int HB, bool NEAR>
__global__ void __launch_bounds__(1 << (LB - 4), 2) ntt_tile_fwd_kernel(const NttArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t limb, tile, grp;
    decode(blockIdx.x, a.groups, 1u << K1, a.l0, limb, tile, grp);
    if (threadIdx.x == 0) { mbar_init(TileSmem<LB>::bar(smem_raw), 1); mbar_fence_init(); }
    __syncthreads();
    tile_fwd_item<LB, K1, HB, NEAR>(a, smem_raw, limb, tile, a.b0 + grp, a.b0 + a.nb, a.groups, 0);
}
template <int LB, int K1, int HB, bool NEAR>
__global__ void __launch_bounds__(1 << (LB - 4), 2) ntt_tile_inv_kernel(const NttArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t limb, tile, grp;
    decode(blockIdx.x, a.groups, 1u << K1, a.l0, limb, tile, grp);
    if (threadIdx.x == 0) { mbar_init(TileSmem<LB>::bar(smem_raw), 1); mbar_fence_init(); }
    __syncthreads();
    tile_inv_item<LB, K1, HB, NEAR>(a, smem_raw, limb, tile, a.b0 + grp, a.b0 + a.nb, a.groups, 0);
}
template <int LB, int K1, int HB, bool NEAR, int V>
__global__ void __launch_bounds__(kRowThreads) ntt_row_fwd_kernel(const NttArgs a) {
    constexpr uint32_t blocks_per_pl = (1u << LB) / (V * kRowThreads);
    uint32_t limb, cb, poly;
    decode(blockIdx.x, a.nb, blocks_per_pl, a.l0, limb, cb, poly);
    row_fwd_item<LB, K1, HB, NEAR, V>(a, limb, a.b0 + poly, cb);
}
template <int LB, int K1, int HB, bool NEAR, int V>
__global__ void __launch_bounds__(kRowThreads) ntt_row_inv_kernel(const NttArgs a) {
    constexpr uint32_t blocks_per_pl = (1u << LB) / (V * kRowThreads);
    uint32_t limb, cb, poly;
    decode(blockIdx.x, a.nb, blocks_per_pl, a.l0, limb, cb, poly);
    row_inv_item<LB, K1, HB, NEAR, V>(a, limb, a.b0 + poly, cb);
}

// ================================================ fused negacyclic product (N <= 4096) ===================================
// out = INTT(NTT(a) . NTT(b)) of one limb-polynomial per CTA, everything in shared memory: replaces the five launches (and the
// two device allocations) of NTTEngine::multiply, /root/reference/src/ntt.cu:49-75, for the sizes where a transform is a single
// tile (BASELINE.json config 1).  The forward tile pass leaves the normalised transform in shared memory in exactly the
// swizzled layout the inverse tile pass starts from, so the two transforms and the pointwise product chain without a copy.
struct MulArgs {
    uint64_t* out; const uint64_t *a, *b;
    const Twiddle *f12, *f3, *i12, *i3;          // staged blocks of both directions, [limbs][1 tile][...]
    const LimbParams* params;
    uint32_t n, limb_count, limb_begin, p3n;
};
template <int LB>
struct MulSmem {
    static constexpr size_t data = (size_t)(1 << LB) * 8, p12 = 256 * sizeof(Twiddle), p3 = (size_t)p3_entries(LB) * sizeof(Twiddle);
    static constexpr size_t total = 2 * data + 2 * (p12 + p3) + 16;
};
template <int LB, int HB, bool NEAR>
__global__ void __launch_bounds__(1 << (LB - 4), 1) negacyclic_mul_kernel(const MulArgs a) {
    using F = TileFwd<LB, HB, NEAR>;
    using I = TileInv<LB, HB, NEAR>;
    using SM = MulSmem<LB>;
    extern __shared__ __align__(128) unsigned char raw[];
    u64* sa = reinterpret_cast<u64*>(raw);
    u64* sb = reinterpret_cast<u64*>(raw + SM::data);
    Twiddle* f12 = reinterpret_cast<Twiddle*>(raw + 2 * SM::data);
    Twiddle* f3 = reinterpret_cast<Twiddle*>(raw + 2 * SM::data + SM::p12);
    Twiddle* i12 = reinterpret_cast<Twiddle*>(raw + 2 * SM::data + SM::p12 + SM::p3);
    Twiddle* i3 = reinterpret_cast<Twiddle*>(raw + 2 * SM::data + 2 * SM::p12 + SM::p3);
    uint64_t* bar = reinterpret_cast<uint64_t*>(raw + 2 * SM::data + 2 * (SM::p12 + SM::p3));
    const uint32_t tid = threadIdx.x;
    const uint32_t limb = blockIdx.x % a.limb_count, pl = a.limb_begin + limb;
    const size_t off = (size_t)blockIdx.x * a.n;                       // [batch][limb_count][n], CTA = (poly, limb)
    if (tid == 0) {
        mbar_init(bar, 1); mbar_fence_init();
        mbar_arrive_expect_tx(bar, (uint32_t)(2 * (SM::p12 + SM::p3)));
        bulk_copy_g2s(f12, a.f12 + (size_t)pl * 256, (uint32_t)SM::p12, bar);
        bulk_copy_g2s(f3, a.f3 + (size_t)pl * a.p3n, (uint32_t)SM::p3, bar);
        bulk_copy_g2s(i12, a.i12 + (size_t)pl * 256, (uint32_t)SM::p12, bar);
        bulk_copy_g2s(i3, a.i3 + (size_t)pl * a.p3n, (uint32_t)SM::p3, bar);
    }
    const LimbParams P = a.params[pl];
    u64 xa[16], xb[16];
    F::phase1_load(tid, a.a + off, xa);
    F::phase1_load(tid, a.b + off, xb);
    __syncthreads();                                                   // barrier initialised before anyone waits on it
    mbar_wait(bar, 0);
    F::template phase1_compute<1>(tid, xa, sa, f12, P);
    F::template phase1_compute<1>(tid, xb, sb, f12, P);
    __syncthreads();
    F::template phase2<1>(tid, sa, f12, P); F::template phase2<1>(tid, sb, f12, P);
    __syncthreads();
    F::template phase3<1>(tid, sa, f3, P); F::template phase3<1>(tid, sb, f3, P);
    __syncthreads();
    constexpr int NT = 1 << (LB - 4);
#pragma unroll
    for (int i = 0; i < 16; i++) { const uint32_t j = tid + i * NT; sa[j] = mul_mod(sa[j], sb[j], P); }     // same layout in both buffers
    __syncthreads();
    I::phase2(tid, sa, i3, P);
    __syncthreads();
    I::phase3(tid, sa, i12, P);
    __syncthreads();
    I::template phase4<true>(tid, a.out + off, sa, i12, P);
}

template <int LB, int HB, bool NEAR>
static int run_mul(fhe_b200_plan* plan, const MulArgs& a, uint32_t ctas, cudaStream_t st) {
    constexpr size_t smem = MulSmem<LB>::total;
    static PerDeviceOnce attr_once;
    if (attr_once.need(plan->device)) {
        FHE_CUDA(cudaFuncSetAttribute(negacyclic_mul_kernel<LB, HB, NEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    negacyclic_mul_kernel<LB, HB, NEAR><<<ctas, 1 << (LB - 4), smem, st>>>(a);
    FHE_LAUNCH_CHECK();
    return 0;
}
template <int HB, bool NEAR>
static int dispatch_mul(fhe_b200_plan* plan, const MulArgs& a, uint32_t ctas, cudaStream_t st) {
    switch (plan->logn) {
        case 9: return run_mul<9, HB, NEAR>(plan, a, ctas, st);
        case 10: return run_mul<10, HB, NEAR>(plan, a, ctas, st);
        case 11: return run_mul<11, HB, NEAR>(plan, a, ctas, st);
        case 12: return run_mul<12, HB, NEAR>(plan, a, ctas, st);
    }
    set_error("fused negacyclic product: N = 2^%u is not a single tile", plan->logn);
    return FHE_B200_EINVAL;
}
// out may alias a or b (every CTA reads its whole inputs before it writes)
int launch_negacyclic_mul_fused(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_a, const uint64_t* d_b, uint32_t batch,
                                uint32_t limb_begin, uint32_t limb_count, cudaStream_t st) {
    MulArgs a;
    a.out = d_out; a.a = d_a; a.b = d_b;
    a.f12 = plan->d_fwd_p12; a.f3 = plan->d_fwd_p3; a.i12 = plan->d_inv_p12; a.i3 = plan->d_inv_p3;
    a.params = plan->d_params; a.n = plan->n; a.limb_count = limb_count; a.limb_begin = limb_begin; a.p3n = (uint32_t)plan->p3_entries;
    const uint32_t ctas = batch * limb_count;
    return plan->near60 ? dispatch_mul<16, true>(plan, a, ctas, st)
         : plan->hb == 16 ? dispatch_mul<16, false>(plan, a, ctas, st)
                          : dispatch_mul<8, false>(plan, a, ctas, st);
}

// ================================================ launch logic =========================================================
template <int LB, int K1, int HB, bool NEAR>
static int run_chunk(fhe_b200_plan* plan, NttArgs a, bool inverse, cudaStream_t st) {
    constexpr int V = (K1 >= 5) ? 1 : 2;
    const int sm_count = plan->sm_count;
    const uint32_t pls = a.nl * a.nb;
    const bool prof = profile_on();
    // tile pass: one CTA per (limb, tile, group); a CTA reuses its staged twiddles for every polynomial of its group.
    // Keep >= 8 CTAs per SM in the grid when the batch allows it, otherwise reuse as much as possible (8 polynomials/CTA).
    const uint32_t lt = a.nl << K1;
    uint32_t groups = (uint32_t)((8u * sm_count + lt - 1) / lt);
    const uint32_t min_groups = (a.nb + 7) / 8;
    if (groups < min_groups) groups = min_groups;
    if (groups > a.nb) groups = a.nb;
    a.groups = groups;
    const uint32_t tile_grid = lt * groups;
    constexpr size_t smem = TileSmem<LB>::total;
    static PerDeviceOnce attr_once;
    if (attr_once.need(plan->device)) {
        FHE_CUDA(cudaFuncSetAttribute(ntt_tile_fwd_kernel<LB, K1, HB, NEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FHE_CUDA(cudaFuncSetAttribute(ntt_tile_inv_kernel<LB, K1, HB, NEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (!inverse) {
        if constexpr (K1 > 0) {
            const uint32_t row_grid = pls * ((1u << LB) / (V * kRowThreads));
            if (prof) profile_begin(2, pls, st);
            ntt_row_fwd_kernel<LB, K1, HB, NEAR, V><<<row_grid, kRowThreads, 0, st>>>(a);
            if (prof) profile_end(st);
            FHE_LAUNCH_CHECK();
        }
        if (prof) profile_begin(0, pls, st);
        ntt_tile_fwd_kernel<LB, K1, HB, NEAR><<<tile_grid, 1 << (LB - 4), smem, st>>>(a);
        if (prof) profile_end(st);
        FHE_LAUNCH_CHECK();
    } else {
        if (prof) profile_begin(1, pls, st);
        ntt_tile_inv_kernel<LB, K1, HB, NEAR><<<tile_grid, 1 << (LB - 4), smem, st>>>(a);
        if (prof) profile_end(st);
        FHE_LAUNCH_CHECK();
        if constexpr (K1 > 0) {
            const uint32_t row_grid = pls * ((1u << LB) / (V * kRowThreads));
            if (prof) profile_begin(3, pls, st);
            ntt_row_inv_kernel<LB, K1, HB, NEAR, V><<<row_grid, kRowThreads, 0, st>>>(a);
            if (prof) profile_end(st);
            FHE_LAUNCH_CHECK();
        }
    }
    return 0;
}

template <int HB, bool NEAR>
static int dispatch(fhe_b200_plan* plan, const NttArgs& a, bool inverse, cudaStream_t st) {
    switch (plan->logn) {
        case 9: return run_chunk<9, 0, HB, NEAR>(plan, a, inverse, st);
        case 10: return run_chunk<10, 0, HB, NEAR>(plan, a, inverse, st);
        case 11: return run_chunk<11, 0, HB, NEAR>(plan, a, inverse, st);
        case 12: return run_chunk<12, 0, HB, NEAR>(plan, a, inverse, st);
        case 13: return run_chunk<12, 1, HB, NEAR>(plan, a, inverse, st);
        case 14: return run_chunk<12, 2, HB, NEAR>(plan, a, inverse, st);
        case 15: return run_chunk<12, 3, HB, NEAR>(plan, a, inverse, st);
        case 16: return run_chunk<12, 4, HB, NEAR>(plan, a, inverse, st);
        case 17: return run_chunk<12, 5, HB, NEAR>(plan, a, inverse, st);
    }
    set_error("unsupported ring degree 2^%u (supported: 2^9 .. 2^17)", plan->logn);
    return FHE_B200_EINVAL;
}

int launch_ntt(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t batch, uint32_t limb_begin,
               uint32_t limb_count, bool inverse, cudaStream_t st) {
    FHE_TRY(check_range(plan, batch, limb_begin, limb_count));
    if (batch == 0 || limb_count == 0) return 0;
    DeviceGuard dev_guard(plan->device);
    NttArgs a;
    a.out = d_out; a.in = d_in;
    a.tw = inverse ? plan->d_inv : plan->d_fwd;
    a.p12 = inverse ? plan->d_inv_p12 : plan->d_fwd_p12;
    a.p3 = inverse ? plan->d_inv_p3 : plan->d_fwd_p3;
    a.tiles = plan->tiles; a.p3n = (uint32_t)plan->p3_entries; a.groups = 1;
    a.params = plan->d_params;
    a.n = plan->n; a.limb_count = limb_count; a.limb_begin = limb_begin;
    const size_t pl_bytes = (size_t)plan->n * sizeof(uint64_t);
    uint32_t nb = batch, nl = limb_count;
    if (plan->logn > 12) {
        // chunking: one limb across many polynomials first (twiddle reuse), several limbs when the batch is small
        const size_t per_chunk = plan->chunk_bytes / pl_bytes ? plan->chunk_bytes / pl_bytes : 1;
        nb = (uint32_t)(batch < per_chunk ? batch : per_chunk);
        const size_t lim = per_chunk / nb ? per_chunk / nb : 1;
        nl = (uint32_t)(limb_count < lim ? limb_count : lim);
        if (nl > 256) nl = 256;
    }
    for (uint32_t l0 = 0; l0 < limb_count; l0 += nl)
        for (uint32_t b0 = 0; b0 < batch; b0 += nb) {
            a.l0 = l0; a.nl = (l0 + nl <= limb_count) ? nl : limb_count - l0;
            a.b0 = b0; a.nb = (b0 + nb <= batch) ? nb : batch - b0;
            if (plan->bal) {
                const int rc = launch_ntt_bal(plan, d_out, d_in, limb_begin, limb_count, a.l0, a.nl, a.b0, a.nb, inverse, st);
                if (rc) return rc;
                continue;
            }
            const int rc = plan->near60 ? dispatch<16, true>(plan, a, inverse, st)
                         : plan->hb == 16 ? dispatch<16, false>(plan, a, inverse, st)
                                          : dispatch<8, false>(plan, a, inverse, st);
            if (rc) return rc;
        }
    return 0;
}

}  // namespace fhe_b200
