// __global__ wrappers and launchers of the batched negacyclic NTT (bodies in ntt_core.cuh).
//
// Work decomposition over a batch laid out [batch][limb_count][N]:
//   tile pass : one CTA of 2^(LB-4) threads per (limb, tile, polynomial group).  The CTA stages the tile's twiddles in
//               shared memory once (two cp.async.bulk copies completing on an mbarrier) and reuses them for every
//               polynomial of its group; CTAs that share (limb, tile) are adjacent in blockIdx (L2 hits).
//   row pass  : one thread per V adjacent columns of a (limb, polynomial); 256-thread CTAs.
// For N > 4096 the two passes are issued chunk by chunk (plan->chunk_bytes of polynomials at a time) so that the
// intermediate written by the first pass is still L2-resident (126 MB on B200) when the second pass reads it:
// HBM then sees one read and one write per limb-transform.
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
    uint32_t groups;             // tile pass: CTAs per (limb, tile); CTA g handles polynomials b0+g, b0+g+groups, ...
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

// 16-byte asynchronous global -> shared copies (LDGSTS) of one tile into the swizzled layout
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
    static constexpr size_t total = data_bytes + p12_bytes + p3_bytes + 16;
};

// One CTA = one (limb, tile): stages that tile's twiddles once with two bulk copies (TMA), then runs the tile pass for
// every polynomial of its group out of shared memory.
template <int LB, int K1, int HB, bool NEAR>
__global__ void __launch_bounds__(1 << (LB - 4), 2) ntt_tile_fwd_kernel(const NttArgs a) {
    using T = TileFwd<LB, HB, NEAR>;
    using SM = TileSmem<LB>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u64* s = reinterpret_cast<u64*>(smem_raw);
    Twiddle* s12 = reinterpret_cast<Twiddle*>(smem_raw + SM::data_bytes);
    Twiddle* s3 = reinterpret_cast<Twiddle*>(smem_raw + SM::data_bytes + SM::p12_bytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + SM::data_bytes + SM::p12_bytes + SM::p3_bytes);
    uint32_t limb, tile, grp;
    decode(blockIdx.x, a.groups, 1u << K1, a.l0, limb, tile, grp);
    const uint32_t pl = a.limb_begin + limb;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar, (uint32_t)(SM::p12_bytes + SM::p3_bytes));
        bulk_copy_g2s(s12, a.p12 + ((size_t)pl * a.tiles + tile) * 256, (uint32_t)SM::p12_bytes, bar);
        bulk_copy_g2s(s3, a.p3 + ((size_t)pl * a.tiles + tile) * a.p3n, (uint32_t)SM::p3_bytes, bar);
    }
    const u64 q = a.params[pl].q;
    constexpr int B0 = fwd_bound_after(1, K1, HB, NEAR);
    const u64* base_in = (K1 > 0 ? a.out : a.in);             // K1 > 0: the row pass already moved the data to `out`
    const size_t limb_off = (size_t)limb * a.n + (size_t)tile * T::NB;
    const size_t poly_stride = (size_t)a.limb_count * a.n;
    uint32_t poly = a.b0 + grp;
    const uint32_t poly_end = a.b0 + a.nb;
    u64 x[16];
    if (poly < poly_end) T::phase1_load(tid, base_in + poly * poly_stride + limb_off, x);
    mbar_wait(bar, 0);                                        // staged twiddles have landed
    for (; poly < poly_end; poly += a.groups) {
        const size_t off = poly * poly_stride + limb_off;
        T::template phase1_compute<B0>(tid, x, s, s12, q);
        __syncthreads();
        T::template phase2<B0>(tid, s, s12, q);
        __syncthreads();
        T::template phase3<B0>(tid, s, s3, q);
        __syncthreads();
        // software pipeline: the next polynomial's global loads fly while this one is copied out
        const uint32_t next = poly + a.groups;
        if (next < poly_end) T::phase1_load(tid, base_in + next * poly_stride + limb_off, x);
        T::phase4(tid, a.out + off, s);
        __syncthreads();
    }
}

template <int LB, int K1, int HB, bool NEAR>
__global__ void __launch_bounds__(1 << (LB - 4), 2) ntt_tile_inv_kernel(const NttArgs a) {
    using T = TileInv<LB, HB, NEAR>;
    using SM = TileSmem<LB>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u64* s = reinterpret_cast<u64*>(smem_raw);
    Twiddle* s12 = reinterpret_cast<Twiddle*>(smem_raw + SM::data_bytes);
    Twiddle* s3 = reinterpret_cast<Twiddle*>(smem_raw + SM::data_bytes + SM::p12_bytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + SM::data_bytes + SM::p12_bytes + SM::p3_bytes);
    uint32_t limb, tile, grp;
    decode(blockIdx.x, a.groups, 1u << K1, a.l0, limb, tile, grp);
    const uint32_t pl = a.limb_begin + limb;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar, (uint32_t)(SM::p12_bytes + SM::p3_bytes));
        bulk_copy_g2s(s12, a.p12 + ((size_t)pl * a.tiles + tile) * 256, (uint32_t)SM::p12_bytes, bar);
        bulk_copy_g2s(s3, a.p3 + ((size_t)pl * a.tiles + tile) * a.p3n, (uint32_t)SM::p3_bytes, bar);
    }
    const LimbParams P = a.params[pl];
    const size_t limb_off = (size_t)limb * a.n + (size_t)tile * T::NB;
    const size_t poly_stride = (size_t)a.limb_count * a.n;
    uint32_t poly = a.b0 + grp;
    const uint32_t poly_end = a.b0 + a.nb;
    if (poly < poly_end) tile_copy_in_async<LB>(tid, a.in + poly * poly_stride + limb_off, s);
    mbar_wait(bar, 0);
    for (; poly < poly_end; poly += a.groups) {
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
        const uint32_t next = poly + a.groups;
        if (next < poly_end) tile_copy_in_async<LB>(tid, a.in + next * poly_stride + limb_off, s);   // overlaps the last 4 stages
        T::template phase4_compute<K1 == 0>(tid, x, a.out + off, s12, P);
    }
}

constexpr int kRowThreads = 256;

template <int LB, int K1, int HB, bool NEAR, int V>
__global__ void __launch_bounds__(kRowThreads) ntt_row_fwd_kernel(const NttArgs a) {
    constexpr uint32_t blocks_per_pl = (1u << LB) / (V * kRowThreads);
    uint32_t limb, cb, poly;
    decode(blockIdx.x, a.nb, blocks_per_pl, a.l0, limb, cb, poly);
    poly += a.b0;
    const size_t off = ((size_t)poly * a.limb_count + limb) * a.n;
    const uint32_t pl = a.limb_begin + limb;
    const uint32_t col = (cb * kRowThreads + threadIdx.x) * V;
    RowPass<K1, V, LB, HB, NEAR>::forward(a.out + off, a.in + off, col, a.tw + (size_t)pl * a.n, a.params[pl].q);
}

template <int LB, int K1, int HB, bool NEAR, int V>
__global__ void __launch_bounds__(kRowThreads) ntt_row_inv_kernel(const NttArgs a) {
    constexpr uint32_t blocks_per_pl = (1u << LB) / (V * kRowThreads);
    uint32_t limb, cb, poly;
    decode(blockIdx.x, a.nb, blocks_per_pl, a.l0, limb, cb, poly);
    poly += a.b0;
    const size_t off = ((size_t)poly * a.limb_count + limb) * a.n;
    const uint32_t pl = a.limb_begin + limb;
    const uint32_t col = (cb * kRowThreads + threadIdx.x) * V;
    constexpr int B0 = TileInv<LB, HB, NEAR>::out_bound();
    // in place on `out`: the tile pass wrote there
    RowPass<K1, V, LB, HB, NEAR>::template inverse<B0>(a.out + off, col, a.tw + (size_t)pl * a.n, a.params[pl]);
}

template <int LB, int K1, int HB, bool NEAR>
static int run_chunk(NttArgs a, bool inverse, int sm_count, cudaStream_t st) {
    constexpr int V = (K1 >= 5) ? 1 : 2;
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
    static bool attr_set = false;
    if (!attr_set) {
        FHE_CUDA(cudaFuncSetAttribute(ntt_tile_fwd_kernel<LB, K1, HB, NEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FHE_CUDA(cudaFuncSetAttribute(ntt_tile_inv_kernel<LB, K1, HB, NEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
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
static int dispatch(uint32_t logn, const NttArgs& a, bool inverse, int sm_count, cudaStream_t st) {
    switch (logn) {
        case 9: return run_chunk<9, 0, HB, NEAR>(a, inverse, sm_count, st);
        case 10: return run_chunk<10, 0, HB, NEAR>(a, inverse, sm_count, st);
        case 11: return run_chunk<11, 0, HB, NEAR>(a, inverse, sm_count, st);
        case 12: return run_chunk<12, 0, HB, NEAR>(a, inverse, sm_count, st);
        case 13: return run_chunk<12, 1, HB, NEAR>(a, inverse, sm_count, st);
        case 14: return run_chunk<12, 2, HB, NEAR>(a, inverse, sm_count, st);
        case 15: return run_chunk<12, 3, HB, NEAR>(a, inverse, sm_count, st);
        case 16: return run_chunk<12, 4, HB, NEAR>(a, inverse, sm_count, st);
        case 17: return run_chunk<12, 5, HB, NEAR>(a, inverse, sm_count, st);
    }
    set_error("unsupported ring degree 2^%u (supported: 2^9 .. 2^17)", logn);
    return FHE_B200_EINVAL;
}

int launch_ntt(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t batch, uint32_t limb_begin,
               uint32_t limb_count, bool inverse, cudaStream_t st) {
    FHE_TRY(check_range(plan, batch, limb_begin, limb_count));
    if (batch == 0 || limb_count == 0) return 0;
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
        // chunk so that (polynomials x limbs) of one chunk stay L2 resident between the two passes;
        // prefer "one limb, many polynomials" (twiddle reuse), widen to several limbs when the batch is small
        const size_t per_chunk = plan->chunk_bytes / pl_bytes ? plan->chunk_bytes / pl_bytes : 1;
        nb = (uint32_t)(batch < per_chunk ? batch : per_chunk);
        const size_t lim = per_chunk / nb ? per_chunk / nb : 1;
        nl = (uint32_t)(limb_count < lim ? limb_count : lim);
    }
    for (uint32_t l0 = 0; l0 < limb_count; l0 += nl)
        for (uint32_t b0 = 0; b0 < batch; b0 += nb) {
            a.l0 = l0; a.nl = (l0 + nl <= limb_count) ? nl : limb_count - l0;
            a.b0 = b0; a.nb = (b0 + nb <= batch) ? nb : batch - b0;
            const int rc = plan->near60 ? dispatch<16, true>(plan->logn, a, inverse, plan->sm_count, st)
                         : plan->hb == 16 ? dispatch<16, false>(plan->logn, a, inverse, plan->sm_count, st)
                                          : dispatch<8, false>(plan->logn, a, inverse, plan->sm_count, st);
            if (rc) return rc;
        }
    return 0;
}

}  // namespace fhe_b200
