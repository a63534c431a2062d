// __global__ wrappers and launcher of the balanced two-pass NTT (bodies and layouts: ntt_bal.cuh), 2^13 <= N <= 2^16.
//
//   bal_a_kernel : pass A / A'.  CTA = 256 threads = one limb, a run of M items (item = one polynomial's block of C columns,
//                  4096 elements; M = 1 by default).  The limb's first 2^KA twiddles are staged once.  36 KiB of shared memory,
//                  64 registers: 4 CTAs per SM.
//   bal_b_kernel : pass B / B'.  Every WARP owns one pair of adjacent tiles of one limb; its 8 KiB twiddle block arrives by one
//                  bulk copy (cp.async.bulk on the warp's own mbarrier) and serves every polynomial of the warp's group.  Only
//                  __syncwarp() between phases.  8 warps per CTA (96 KiB), 2 CTAs per SM.
// Forward: A reads `in`, writes `out`; B runs in place on `out`.  Inverse: B' reads `in`, writes `out`; A' in place on `out`.
#include "common.cuh"
#include "ntt_bal.cuh"
#include "tma.cuh"

namespace fhe_b200 {

struct BalArgs {
    uint64_t* out;
    const uint64_t* in;
    const Twiddle* tw;           // [limbs][n] table of the direction in use (pass A reads its first 2^KA entries)
    const Twiddle* blocks;       // [limbs][n/512][512] staged blocks of pass B
    const LimbParams* params;    // [limbs]
    uint32_t n, limb_count, limb_begin;
    uint32_t l0, nl, b0, nb;     // chunk: buffer limbs [l0, l0+nl), polynomials [b0, b0+nb)
    uint32_t m_items, ctas_per_limb;   // pass A: items per CTA, CTAs per limb
    uint32_t groups;             // pass B: warps per (limb, tile pair); warp g handles polynomials b0+g, b0+g+groups, ...
    size_t in_poly_stride;       // elements between polynomials of `in` (the first pass of a transform reads `in`); out: limb_count * n
};

// forward transforms launched while this is set leave their results lazy (below 8q, see BalB::fwd_phase2); set by ScopedLazyForward
thread_local bool g_fwd_lazy_out = false;

constexpr size_t kBalASmem = 4096 * sizeof(u64) + 256 * sizeof(Twiddle);
#ifndef FHE_BAL_B_MINBLOCKS
#define FHE_BAL_B_MINBLOCKS 2
#endif
#ifndef FHE_BAL_B_PAIRS
#define FHE_BAL_B_PAIRS 4
#endif
#ifndef FHE_BAL_B_GROUPS
#define FHE_BAL_B_GROUPS 2
#endif
constexpr int kBalBPairs = FHE_BAL_B_PAIRS, kBalBGroups = FHE_BAL_B_GROUPS, kBalBWarps = kBalBPairs * kBalBGroups, kBalBMinBlocks = FHE_BAL_B_MINBLOCKS;
// per warp: the 4 KiB exchange buffer and a 4 KiB staging buffer that the next polynomial's tile pair is bulk-copied into while this one is
// transformed, with two mbarriers; per pair: the 8 KiB twiddle block and its mbarrier
constexpr size_t kBalBSmem = kBalBWarps * (2 * 512 * sizeof(u64) + 16) + kBalBPairs * (512 * sizeof(Twiddle) + 16);

template <int KA, int HB, bool NEAR, bool INV>
__global__ void __launch_bounds__(256, 4) bal_a_kernel(const BalArgs a) {
    using A = BalA<KA, HB, NEAR>;
    extern __shared__ __align__(128) unsigned char raw[];
    u64* sbuf = reinterpret_cast<u64*>(raw);
    Twiddle* stw = reinterpret_cast<Twiddle*>(raw + 4096 * sizeof(u64));
    const uint32_t tid = threadIdx.x;
    const uint32_t limb = a.l0 + blockIdx.x / a.ctas_per_limb, chunk = blockIdx.x % a.ctas_per_limb;
    const uint32_t pl = a.limb_begin + limb;
    if (tid < (1u << KA)) {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(a.tw + (size_t)pl * a.n + tid));
        *reinterpret_cast<ulonglong2*>(stw + tid) = v;
    }
    const LimbParams P = a.params[pl];
    const uint32_t items = a.nb * A::CB;
    uint32_t it = chunk * a.m_items;
    const uint32_t end = min(items, it + a.m_items);
    constexpr int BIN = BalB<HB, NEAR>::inv_out_bound();
    __syncthreads();
#pragma unroll 1
    for (; it < end; it++) {
        const uint32_t poly = a.b0 + it / A::CB, cb = it % A::CB;
        const size_t off = ((size_t)poly * a.limb_count + limb) * a.n + (size_t)cb * A::C;
        if (!INV) {
            A::fwd_round1(tid, a.in + ((size_t)poly * a.in_poly_stride + (size_t)limb * a.n + (size_t)cb * A::C), sbuf, stw, P);
            __syncthreads();
            A::fwd_round2(tid, a.out + off, sbuf, stw, P);
        } else {
            A::template inv_round2<BIN>(tid, a.out + off, sbuf, stw, P);
            __syncthreads();
            A::template inv_round1<BIN>(tid, a.out + off, sbuf, stw, P);
        }
        if (it + 1 < end) __syncthreads();              // (one item per CTA is the default: no second barrier then)
    }
}

// pass A' with scattered outputs (BalScatter, common.cuh): reads a.out (where pass B' left its result), writes block by block
template <int KA, int HB, bool NEAR>
__global__ void __launch_bounds__(256, 4) bal_a_scatter_kernel(const BalArgs a, const BalScatter sc) {
    using A = BalA<KA, HB, NEAR>;
    extern __shared__ __align__(128) unsigned char raw[];
    u64* sbuf = reinterpret_cast<u64*>(raw);
    Twiddle* stw = reinterpret_cast<Twiddle*>(raw + 4096 * sizeof(u64));
    const uint32_t tid = threadIdx.x;
    const uint32_t limb = a.l0 + blockIdx.x / a.ctas_per_limb, chunk = blockIdx.x % a.ctas_per_limb;
    const uint32_t pl = a.limb_begin + limb;
    if (tid < (1u << KA)) {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(a.tw + (size_t)pl * a.n + tid));
        *reinterpret_cast<ulonglong2*>(stw + tid) = v;
    }
    const LimbParams P = a.params[pl];
    const uint32_t items = a.nb * A::CB;
    uint32_t it = chunk * a.m_items;
    const uint32_t end = min(items, it + a.m_items);
    constexpr int BIN = BalB<HB, NEAR>::inv_out_bound();
    const uint32_t log_rpb = KA - sc.log_blocks;                 // rows per coefficient block = 2^KA / blocks
    __syncthreads();
#pragma unroll 1
    for (; it < end; it++) {
        const uint32_t poly = a.b0 + it / A::CB, cb = it % A::CB;
        const size_t off = ((size_t)poly * a.limb_count + limb) * a.n + (size_t)cb * A::C;
        A::template inv_round2<BIN>(tid, a.out + off, sbuf, stw, P);
        __syncthreads();
        const size_t row0 = ((size_t)poly * sc.limbs_total + sc.limb_off + limb) * sc.nc + (size_t)cb * A::C;
        auto put = [&](u32 row, u32 col, u64 v) {
            const u32 h = row >> log_rpb, rr = row & ((1u << log_rpb) - 1);
            reinterpret_cast<u64*>(sc.base[h])[row0 + ((size_t)rr << 8) + col] = v;
        };
        A::template inv_round1_put<BIN>(tid, put, sbuf, stw, P);
        if (it + 1 < end) __syncthreads();
    }
}

// CTA = kBalBWarps warps = kBalBPairs tile pairs x kBalBGroups polynomial groups: the warps of one pair share its staged
// twiddle block (8 KiB): 4 x 8 KiB of twiddles + 8 x 4 KiB of exchange buffers = 64 KiB per CTA.  Two CTAs (16 warps, 128 registers) per
// SM is the measured optimum at config 3: 0.645 / 0.611 ms per forward / inverse pass, against 0.666 / 0.618 with three CTAs at 80
// registers, 0.655 / 0.599 with five 128-thread CTAs at 96, 0.717 / 0.657 with four CTAs at 64 (spills).
template <int KA, int HB, bool NEAR, bool INV, bool LAZY = false>
__global__ void __launch_bounds__(32 * kBalBWarps, kBalBMinBlocks) bal_b_kernel(const BalArgs a) {
    using B = BalB<HB, NEAR>;
    extern __shared__ __align__(128) unsigned char raw[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t pl_local = warp % kBalBPairs, gl = warp / kBalBPairs;       // pair and group inside the CTA
    u64* s = reinterpret_cast<u64*>(raw) + warp * 512;                         // exchange buffer
    u64* stg = reinterpret_cast<u64*>(raw) + (kBalBWarps + warp) * 512;        // staging buffer (bulk copies land here)
    unsigned char* after = raw + 2 * kBalBWarps * 512 * sizeof(u64);
    Twiddle* sb = reinterpret_cast<Twiddle*>(after) + pl_local * 512;
    uint64_t* bar = reinterpret_cast<uint64_t*>(after + kBalBPairs * 512 * sizeof(Twiddle)) + pl_local * 2;
    uint64_t* wbar = reinterpret_cast<uint64_t*>(after + kBalBPairs * (512 * sizeof(Twiddle) + 16)) + warp * 2;   // this warp's two barriers
    constexpr uint32_t pairs = 1u << (KA - 1), pair_blocks = pairs / kBalBPairs;
    const uint32_t gblocks = (a.groups + kBalBGroups - 1) / kBalBGroups;
    const uint32_t pb = blockIdx.x % pair_blocks, r = blockIdx.x / pair_blocks, gb = r % gblocks, limb = a.l0 + r / gblocks;
    const uint32_t pair = pb * kBalBPairs + pl_local, grp = gb * kBalBGroups + gl;
    const uint32_t pl = a.limb_begin + limb;
    if (threadIdx.x < kBalBPairs) mbar_init(reinterpret_cast<uint64_t*>(after + kBalBPairs * 512 * sizeof(Twiddle)) + threadIdx.x * 2, 1);
    if (lane < 2) mbar_init(wbar + lane, 1);
    if (threadIdx.x == 0) mbar_fence_init();
    __syncthreads();
    if (gl == 0 && lane == 0) {
        mbar_arrive_expect_tx(bar, 512 * sizeof(Twiddle));
        bulk_copy_g2s(sb, a.blocks + ((size_t)pl * pairs + pair) * 512, 512 * sizeof(Twiddle), bar);
    }
    if (grp >= a.groups) return;                       // odd group count: the second half of the last group block has no work
    const LimbParams P = a.params[pl];
    const size_t limb_off = (size_t)limb * a.n + (size_t)pair * 512;
    const size_t poly_stride = (size_t)a.limb_count * a.n;
    constexpr int B0 = BalA<KA, HB, NEAR>::fwd_out_bound();
    const uint32_t first = a.b0 + grp, end = a.b0 + a.nb;
    // A polynomial's tile pair is 4 KiB of contiguous global memory: one bulk copy (TMA) brings it into shared memory while the warp
    // transforms the previous one -- the load latency costs no registers and no issue slots.  (Loading it early into registers was
    // measured slower: 0.701 ms against 0.646 ms per forward pass at config 3.)
    auto fetch = [&](u64* dst, uint64_t* wb, const u64* src) {      // called by the whole warp after its last read of dst
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_expect_tx(wb, 512 * sizeof(u64));
            bulk_copy_g2s(dst, src, 512 * sizeof(u64), wb);
        }
    };
    if (!INV) {
        // forward: round 1 takes its sixteen values out of the staging buffer, which is then free for the next polynomial
        if (first < end) fetch(stg, wbar, a.out + first * poly_stride + limb_off);
        mbar_wait(bar, 0);
        uint32_t it = 0;
#pragma unroll 1
        for (uint32_t poly = first; poly < end; poly += a.groups, it++) {
            const size_t off = poly * poly_stride + limb_off;
            u64 x[16];
            mbar_wait(wbar, it & 1);
            B::fwd_load_staged(lane, stg, x);
            const size_t nxt = (size_t)poly + a.groups;
            if (nxt < end) fetch(stg, wbar, a.out + nxt * poly_stride + limb_off);
            B::template fwd_phase1<B0>(lane, x, s, sb, P);
            __syncwarp();
            B::template fwd_phase2<B0, LAZY>(lane, s, sb, P);
            __syncwarp();
            B::fwd_phase3(lane, a.out + off, s);
            __syncwarp();
        }
    } else {
        // inverse: one staging buffer; the re-layout into the exchange buffer frees it for the next polynomial
        if (first < end) fetch(stg, wbar, a.in + (first * a.in_poly_stride + limb_off));
        mbar_wait(bar, 0);
        uint32_t it = 0;
#pragma unroll 1
        for (uint32_t poly = first; poly < end; poly += a.groups, it++) {
            const size_t off = poly * poly_stride + limb_off;
            mbar_wait(wbar, it & 1);
            B::inv_phase1_staged(lane, stg, s);
            const size_t nxt = (size_t)poly + a.groups;
            if (nxt < end) fetch(stg, wbar, a.in + (nxt * a.in_poly_stride + limb_off));
            else __syncwarp();
            B::inv_phase2(lane, s, sb, P);
            __syncwarp();
            B::inv_phase3(lane, a.out + off, s, sb, P);
            __syncwarp();
        }
    }
}

#ifndef FHE_BAL_EXPERIMENT          // (tools/sass/balexp.cu instantiates single kernels for SASS accounting)
int launch_ntt_bal_passes(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t limb_begin, uint32_t limb_count,
                          uint32_t l0, uint32_t nl, uint32_t b0, uint32_t nb, bool inverse, cudaStream_t st, const BalScatter* scatter,
                          bool only_a);
extern thread_local size_t g_in_poly_stride_fwd;
template <int KA, int HB, bool NEAR>
static int run_bal_chunk(fhe_b200_plan* plan, BalArgs a, bool inverse, cudaStream_t st, const BalScatter* scatter, bool only_a) {
    using A = BalA<KA, HB, NEAR>;
    static PerDeviceOnce attr_once;
    if (attr_once.need(plan->device)) {
        FHE_CUDA(cudaFuncSetAttribute(bal_a_kernel<KA, HB, NEAR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBalASmem));
        FHE_CUDA(cudaFuncSetAttribute(bal_a_kernel<KA, HB, NEAR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBalASmem));
        FHE_CUDA(cudaFuncSetAttribute(bal_a_scatter_kernel<KA, HB, NEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBalASmem));
        FHE_CUDA(cudaFuncSetAttribute(bal_b_kernel<KA, HB, NEAR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBalBSmem));
        FHE_CUDA(cudaFuncSetAttribute(bal_b_kernel<KA, HB, NEAR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBalBSmem));
        FHE_CUDA(cudaFuncSetAttribute(bal_b_kernel<KA, HB, NEAR, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBalBSmem));
    }
    const uint32_t sms = (uint32_t)plan->sm_count;
    const uint64_t pls = (uint64_t)a.nl * a.nb;
    // pass A: one item per CTA (measured at config 3: 1 item 0.590 ms, 2 items 0.653 ms, 8 items 0.684 ms per forward pass --
    // the hardware CTA scheduler overlaps the load, compute and store phases of neighbouring items better than a loop does)
    const uint32_t items_per_limb = a.nb * A::CB;
    uint32_t m = 1;
    const int env_m = getenv("FHE_B200_BAL_M") ? atoi(getenv("FHE_B200_BAL_M")) : 0;          // experiments
    if (env_m > 0) m = (uint32_t)env_m;
    a.m_items = m;
    a.ctas_per_limb = (items_per_limb + m - 1) / m;
    const uint32_t grid_a = a.nl * a.ctas_per_limb;
    // pass B: one warp per (limb, tile pair, group); a CTA covers kBalBPairs pairs x kBalBGroups groups.  Few groups = many
    // polynomials per staged twiddle block (measured at config 3: 2 groups 0.646 ms, 3 groups 0.669 ms, 8 groups 0.657 ms,
    // 1 group 0.740 ms (under two waves)): aim at three waves of resident warps
    constexpr uint32_t pairs = 1u << (KA - 1);
    const uint32_t lp = a.nl * pairs;
    uint32_t groups = (4u * kBalBMinBlocks * sms * kBalBWarps + lp - 1) / lp;
    const int env_g = getenv("FHE_B200_BAL_GROUPS") ? atoi(getenv("FHE_B200_BAL_GROUPS")) : 0;
    if (env_g > 0) groups = (uint32_t)env_g;
    groups = (groups + kBalBGroups - 1) / kBalBGroups * kBalBGroups;       // whole CTAs (a CTA spans kBalBGroups groups)
    groups = groups < 1 ? 1 : (groups > a.nb ? a.nb : groups);
    a.groups = groups;
    const uint32_t grid_b = a.nl * (pairs / kBalBPairs) * ((groups + kBalBGroups - 1) / kBalBGroups);
    const bool prof = profile_on();
    if (!inverse) {
        if (prof) profile_begin(2, pls, st);
        bal_a_kernel<KA, HB, NEAR, false><<<grid_a, 256, kBalASmem, st>>>(a);
        if (prof) profile_end(st);
        FHE_LAUNCH_CHECK();
        if (only_a) return 0;
        if (prof) profile_begin(0, pls, st);
        if (g_fwd_lazy_out) bal_b_kernel<KA, HB, NEAR, false, true><<<grid_b, 32 * kBalBWarps, kBalBSmem, st>>>(a);
        else bal_b_kernel<KA, HB, NEAR, false><<<grid_b, 32 * kBalBWarps, kBalBSmem, st>>>(a);
        if (prof) profile_end(st);
        FHE_LAUNCH_CHECK();
    } else {
        if (!only_a) {
            if (prof) profile_begin(1, pls, st);
            bal_b_kernel<KA, HB, NEAR, true><<<grid_b, 32 * kBalBWarps, kBalBSmem, st>>>(a);
            if (prof) profile_end(st);
            FHE_LAUNCH_CHECK();
        }
        if (prof) profile_begin(3, pls, st);
        if (scatter) bal_a_scatter_kernel<KA, HB, NEAR><<<grid_a, 256, kBalASmem, st>>>(a, *scatter);
        else bal_a_kernel<KA, HB, NEAR, true><<<grid_a, 256, kBalASmem, st>>>(a);
        if (prof) profile_end(st);
        FHE_LAUNCH_CHECK();
    }
    return 0;
}

template <int HB, bool NEAR>
static int dispatch_bal(fhe_b200_plan* plan, const BalArgs& a, bool inverse, cudaStream_t st, const BalScatter* scatter, bool only_a) {
    switch (plan->logn) {
        case 13: return run_bal_chunk<5, HB, NEAR>(plan, a, inverse, st, scatter, only_a);
        case 14: return run_bal_chunk<6, HB, NEAR>(plan, a, inverse, st, scatter, only_a);
        case 15: return run_bal_chunk<7, HB, NEAR>(plan, a, inverse, st, scatter, only_a);
        case 16: return run_bal_chunk<8, HB, NEAR>(plan, a, inverse, st, scatter, only_a);
    }
    set_error("balanced NTT: unsupported ring degree 2^%u", plan->logn);
    return FHE_B200_EINVAL;
}

// chunk = buffer limbs [l0, l0+nl) x polynomials [b0, b0+nb) of a [batch][limb_count][n] buffer
// polynomial stride of the INPUT of the transform being launched (elements; 0 = dense).  Set by launch_ntt_strided_in around its call.
thread_local size_t g_in_poly_stride_fwd = 0;
#define g_in_poly_stride g_in_poly_stride_fwd

// out-of-place transform whose input polynomials are `in_poly_stride` elements apart (e.g. one component of interleaved
// ciphertexts): saves the gather copy.  Balanced two-pass sizes only; out is dense [batch][limb_count][N].
int launch_ntt_strided_in(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, size_t in_poly_stride, uint32_t batch,
                          uint32_t limb_begin, uint32_t limb_count, bool inverse, cudaStream_t st) {
    FHE_TRY(check_range(plan, batch, limb_begin, limb_count));
    FHE_REQUIRE(plan->bal && d_out != d_in, "strided transform: needs the balanced two-pass NTT and distinct buffers");
    if (batch == 0 || limb_count == 0) return 0;
    DeviceGuard dev_guard(plan->device);
    g_in_poly_stride = in_poly_stride;
    int rc = 0;
    for (uint32_t l0 = 0; l0 < limb_count && !rc; l0 += 256) {
        const uint32_t nl = l0 + 256 <= limb_count ? 256 : limb_count - l0;
        rc = launch_ntt_bal_passes(plan, d_out, d_in, limb_begin, limb_count, l0, nl, 0, batch, inverse, st, nullptr, false);
    }
    g_in_poly_stride = 0;
    return rc;
}

int launch_ntt_bal(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t limb_begin, uint32_t limb_count,
                   uint32_t l0, uint32_t nl, uint32_t b0, uint32_t nb, bool inverse, cudaStream_t st, const BalScatter* scatter) {
    return launch_ntt_bal_passes(plan, d_out, d_in, limb_begin, limb_count, l0, nl, b0, nb, inverse, st, scatter, false);
}

int launch_ntt_bal_passes(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t limb_begin, uint32_t limb_count,
                          uint32_t l0, uint32_t nl, uint32_t b0, uint32_t nb, bool inverse, cudaStream_t st, const BalScatter* scatter,
                          bool only_a) {
    BalArgs a;
    a.out = d_out; a.in = d_in;
    a.tw = inverse ? plan->d_inv : plan->d_fwd;
    a.blocks = inverse ? plan->d_inv_bal : plan->d_fwd_bal;
    a.params = plan->d_params;
    a.n = plan->n; a.limb_count = limb_count; a.limb_begin = limb_begin;
    a.l0 = l0; a.nl = nl; a.b0 = b0; a.nb = nb;
    a.m_items = 1; a.ctas_per_limb = 1; a.groups = 1;
    a.in_poly_stride = g_in_poly_stride ? g_in_poly_stride : (size_t)limb_count * plan->n;
    return plan->near60 ? dispatch_bal<16, true>(plan, a, inverse, st, scatter, only_a)
         : plan->hb == 16 ? dispatch_bal<16, false>(plan, a, inverse, st, scatter, only_a)
                          : dispatch_bal<8, false>(plan, a, inverse, st, scatter, only_a);
}

// the column pass alone (the fused tile kernels of ntt_fused.cu do the rest): forward in -> out; inverse in place on d_out
// (or scattered).  Limbs in groups of at most 256 (grid size), all polynomials in one launch.
int launch_ntt_pass_a(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t batch, uint32_t limb_begin, uint32_t limb_count,
                      bool inverse, cudaStream_t st, const BalScatter* scatter) {
    FHE_TRY(check_range(plan, batch, limb_begin, limb_count));
    FHE_REQUIRE(plan->bal, "column pass: needs the balanced two-pass NTT (2^13 <= N <= 2^16)");
    if (batch == 0 || limb_count == 0) return 0;
    DeviceGuard dev_guard(plan->device);
    for (uint32_t l0 = 0; l0 < limb_count; l0 += 256) {
        const uint32_t nl = l0 + 256 <= limb_count ? 256 : limb_count - l0;
        FHE_TRY(launch_ntt_bal_passes(plan, d_out, d_in, limb_begin, limb_count, l0, nl, 0, batch, inverse, st, scatter, true));
    }
    return 0;
}

int launch_ntt_inverse_scatter(fhe_b200_plan* plan, uint64_t* d_buf, uint32_t batch, uint32_t limb_begin, uint32_t limb_count,
                               const BalScatter& scatter, cudaStream_t st) {
    FHE_TRY(check_range(plan, batch, limb_begin, limb_count));
    FHE_REQUIRE(plan->bal, "scatter transform: needs the balanced two-pass NTT (2^13 <= N <= 2^16)");
    FHE_REQUIRE(scatter.log_blocks <= 4 && scatter.log_blocks + 8 <= plan->logn && (scatter.nc << scatter.log_blocks) == plan->n,
                "scatter transform: bad coefficient blocks");
    if (batch == 0 || limb_count == 0) return 0;
    DeviceGuard dev_guard(plan->device);
    // limbs in groups of at most 256 (grid size); all polynomials in one launch
    for (uint32_t l0 = 0; l0 < limb_count; l0 += 256) {
        const uint32_t nl = l0 + 256 <= limb_count ? 256 : limb_count - l0;
        FHE_TRY(launch_ntt_bal(plan, d_buf, d_buf, limb_begin, limb_count, l0, nl, 0, batch, true, st, &scatter));
    }
    return 0;
}
#endif

}  // namespace fhe_b200
