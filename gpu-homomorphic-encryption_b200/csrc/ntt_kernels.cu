// __global__ wrappers and launchers of the batched negacyclic NTT (bodies in ntt_core.cuh).
//
// Work decomposition over a batch laid out [batch][limb_count][N]:
//   tile pass : one CTA of 2^(LB-4) threads per (limb, tile, polynomial); CTAs that share (limb, tile) -- hence the
//               same twiddles -- are adjacent in blockIdx so the twiddle lines are hot in L2 when their
//               neighbours ask for them.
//   row pass  : one thread per V adjacent columns of a (limb, polynomial); 256-thread CTAs.
// For N > 4096 the two passes are issued chunk by chunk (plan->chunk_bytes of polynomials at a time) so that the
// intermediate written by the first pass is still L2-resident (126 MB on B200) when the second pass reads it:
// HBM then sees one read and one write per limb-transform.
#include "common.cuh"
#include "ntt_core.cuh"

namespace fhe_b200 {

struct NttArgs {
    uint64_t* out;
    const uint64_t* in;
    const Twiddle* tw;           // [limbs][n] table of the direction in use
    const LimbParams* params;    // [limbs]
    uint32_t n;                  // ring degree
    uint32_t limb_count;         // limbs per polynomial in the buffer
    uint32_t limb_begin;         // first plan limb the buffer's limb 0 maps to
    uint32_t l0, nl;             // chunk: buffer limbs [l0, l0+nl)
    uint32_t b0, nb;             // chunk: polynomials [b0, b0+nb)
};

// (limb, tile/column-block, polynomial) from a linear CTA index with the polynomial fastest
__device__ __forceinline__ void decode(const NttArgs& a, uint32_t i, uint32_t units, uint32_t& limb, uint32_t& unit,
                                       uint32_t& poly) {
    poly = a.b0 + i % a.nb;
    const uint32_t r = i / a.nb;
    unit = r % units;
    limb = a.l0 + r / units;
}

template <int LB, int K1, int HB>
__global__ void __launch_bounds__(1 << (LB - 4)) ntt_tile_fwd_kernel(const NttArgs a) {
    using T = TileFwd<LB, HB>;
    __shared__ __align__(16) u64 s[T::NB];
    uint32_t limb, tile, poly;
    decode(a, blockIdx.x, 1u << K1, limb, tile, poly);
    const size_t off = ((size_t)poly * a.limb_count + limb) * a.n + (size_t)tile * T::NB;
    const uint32_t pl = a.limb_begin + limb;
    const Twiddle* tw = a.tw + (size_t)pl * a.n;
    const u64 q = a.params[pl].q;
    const uint32_t root = (1u << K1) + tile;
    constexpr int B0 = fwd_bound_after(1, K1, HB);
    // K1 > 0: the row pass already moved the data to `out`
    const u64* src = (K1 > 0 ? a.out : a.in) + off;
    T::template phase1<B0>(threadIdx.x, src, s, tw, root, q);
    __syncthreads();
    T::template phase2<B0>(threadIdx.x, s, tw, root, q);
    __syncthreads();
    T::template phase3<B0>(threadIdx.x, s, tw, root, q);
    __syncthreads();
    T::phase4(threadIdx.x, a.out + off, s);
}

template <int LB, int K1, int HB>
__global__ void __launch_bounds__(1 << (LB - 4)) ntt_tile_inv_kernel(const NttArgs a) {
    using T = TileInv<LB, HB>;
    __shared__ __align__(16) u64 s[T::NB];
    uint32_t limb, tile, poly;
    decode(a, blockIdx.x, 1u << K1, limb, tile, poly);
    const size_t off = ((size_t)poly * a.limb_count + limb) * a.n + (size_t)tile * T::NB;
    const uint32_t pl = a.limb_begin + limb;
    const Twiddle* tw = a.tw + (size_t)pl * a.n;
    const LimbParams P = a.params[pl];
    const uint32_t root = (1u << K1) + tile;
    T::phase1(threadIdx.x, a.in + off, s);
    __syncthreads();
    T::phase2(threadIdx.x, s, tw, root, P);
    __syncthreads();
    T::phase3(threadIdx.x, s, tw, root, P);
    __syncthreads();
    T::template phase4<K1 == 0>(threadIdx.x, a.out + off, s, tw, root, P);
}

constexpr int kRowThreads = 256;

template <int LB, int K1, int HB, int V>
__global__ void __launch_bounds__(kRowThreads) ntt_row_fwd_kernel(const NttArgs a) {
    constexpr uint32_t blocks_per_pl = (1u << LB) / (V * kRowThreads);
    uint32_t limb, cb, poly;
    decode(a, blockIdx.x, blocks_per_pl, limb, cb, poly);
    const size_t off = ((size_t)poly * a.limb_count + limb) * a.n;
    const uint32_t pl = a.limb_begin + limb;
    const uint32_t col = (cb * kRowThreads + threadIdx.x) * V;
    RowPass<K1, V, LB, HB>::forward(a.out + off, a.in + off, col, a.tw + (size_t)pl * a.n, a.params[pl].q);
}

template <int LB, int K1, int HB, int V>
__global__ void __launch_bounds__(kRowThreads) ntt_row_inv_kernel(const NttArgs a) {
    constexpr uint32_t blocks_per_pl = (1u << LB) / (V * kRowThreads);
    uint32_t limb, cb, poly;
    decode(a, blockIdx.x, blocks_per_pl, limb, cb, poly);
    const size_t off = ((size_t)poly * a.limb_count + limb) * a.n;
    const uint32_t pl = a.limb_begin + limb;
    const uint32_t col = (cb * kRowThreads + threadIdx.x) * V;
    constexpr int B0 = TileInv<LB, HB>::out_bound();
    // in place on `out`: the tile pass wrote there
    RowPass<K1, V, LB, HB>::template inverse<B0>(a.out + off, col, a.tw + (size_t)pl * a.n, a.params[pl]);
}

template <int LB, int K1, int HB>
static int run_chunk(const NttArgs& a, bool inverse, cudaStream_t st) {
    constexpr int V = (K1 >= 5) ? 1 : 2;
    const uint32_t pls = a.nl * a.nb;
    const uint32_t tile_grid = pls << K1;
    const bool prof = profile_on();
    if (!inverse) {
        if constexpr (K1 > 0) {
            const uint32_t row_grid = pls * ((1u << LB) / (V * kRowThreads));
            if (prof) profile_begin(2, pls, st);
            ntt_row_fwd_kernel<LB, K1, HB, V><<<row_grid, kRowThreads, 0, st>>>(a);
            if (prof) profile_end(st);
            FHE_LAUNCH_CHECK();
        }
        if (prof) profile_begin(0, pls, st);
        ntt_tile_fwd_kernel<LB, K1, HB><<<tile_grid, 1 << (LB - 4), 0, st>>>(a);
        if (prof) profile_end(st);
        FHE_LAUNCH_CHECK();
    } else {
        if (prof) profile_begin(1, pls, st);
        ntt_tile_inv_kernel<LB, K1, HB><<<tile_grid, 1 << (LB - 4), 0, st>>>(a);
        if (prof) profile_end(st);
        FHE_LAUNCH_CHECK();
        if constexpr (K1 > 0) {
            const uint32_t row_grid = pls * ((1u << LB) / (V * kRowThreads));
            if (prof) profile_begin(3, pls, st);
            ntt_row_inv_kernel<LB, K1, HB, V><<<row_grid, kRowThreads, 0, st>>>(a);
            if (prof) profile_end(st);
            FHE_LAUNCH_CHECK();
        }
    }
    return 0;
}

template <int HB>
static int dispatch(uint32_t logn, const NttArgs& a, bool inverse, cudaStream_t st) {
    switch (logn) {
        case 9: return run_chunk<9, 0, HB>(a, inverse, st);
        case 10: return run_chunk<10, 0, HB>(a, inverse, st);
        case 11: return run_chunk<11, 0, HB>(a, inverse, st);
        case 12: return run_chunk<12, 0, HB>(a, inverse, st);
        case 13: return run_chunk<12, 1, HB>(a, inverse, st);
        case 14: return run_chunk<12, 2, HB>(a, inverse, st);
        case 15: return run_chunk<12, 3, HB>(a, inverse, st);
        case 16: return run_chunk<12, 4, HB>(a, inverse, st);
        case 17: return run_chunk<12, 5, HB>(a, inverse, st);
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
            const int rc = plan->hb == 16 ? dispatch<16>(plan->logn, a, inverse, st) : dispatch<8>(plan->logn, a, inverse, st);
            if (rc) return rc;
        }
    return 0;
}

}  // namespace fhe_b200
