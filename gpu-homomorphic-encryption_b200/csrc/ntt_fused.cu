// Fused tile passes of the BFV multiply: the pointwise work that sits between a forward and an inverse transform runs inside the
// tile pass, on the 256-element tiles both passes share, instead of as two more launches and a round trip through HBM.
//
//   tensor : (a0, a1, b0, b1) after the column pass  ->  tile pass x4  ->  d0 = a0 b0, d1 = a0 b1 + a1 b0, d2 = a1 b1
//            ->  inverse tile pass x3  ->  (d0, d1, d2) ready for the inverse column pass
//            replaces bal_b<fwd> + tensor_kernel + bal_b<inv>   (the products of FHEContext::multiply, /root/reference/src/fhe.cu:199-224,
//            each an NTTEngine::multiply there, src/ntt.cu:49-75)
//   inner  : dnum key-switch digits after the column pass  ->  tile pass x dnum  ->  acc_c = sum_d digit_d * key[d][c], c = 0, 1
//            ->  inverse tile pass x2                        replaces bal_b<fwd> + ks_inner_kernel + bal_b<inv>
//            (the relinearisation the reference stubs out, src/fhe.cu:226-235; intent docs/ARCHITECTURE.md:319-326)
//
// A WARP owns one pair of adjacent tiles of one limb, as in bal_b_kernel (ntt_bal.cu), and walks the ciphertexts of its group.
// It has three 4 KiB buffers: a transformed operand stays in the buffer it was exchanged through (round 2 reads and writes the
// lane's OWN row), the fourth operand of the tensor product stays in registers, the products overwrite the operands row by row,
// and the inverse rounds start from those rows -- no block barrier and, between the forward and inverse halves, not even a warp
// barrier.  The pair's forward and inverse twiddle blocks (2 x 8 KiB, one bulk copy each) are shared by the warps of the CTA that
// work on that pair.  PAIRS x GROUPS warps per CTA: 4 x 2 (160 KiB, batch >= 2) or 8 x 1 (224 KiB, one ciphertext); one CTA per SM.
#include "common.cuh"
#include "ntt_bal.cuh"
#include "tma.cuh"

namespace fhe_b200 {

struct FusedArgs {
    uint64_t* out; const uint64_t* in;
    const Twiddle *blocks_f, *blocks_i;      // [limbs][n/512][512] staged blocks of the tile pass, both directions
    const LimbParams* params;
    uint32_t n, limb_begin, limb_count;      // limb l of the buffers is plan limb limb_begin + l, l < limb_count
    uint32_t nb, groups;                     // ciphertexts; warps per (limb, tile pair): warp g handles ciphertexts g, g + groups, ...
    size_t in_plane[4], in_poly;             // element offsets: operand p of ciphertext b, limb l at in + in_plane[p] + b * in_poly + l * n
    size_t out_plane[3], out_poly;
    const uint64_t* key; size_t key_poly;    // inner product: key polynomial (d, c) at key + (2 d + c) * key_poly, limb l at + l * n
    uint32_t dnum, square;
};

// NBUF buffers of 4 KiB per warp: 4 = every operand of the tensor product has its own (batch >= 2: 8 warps, 192 KiB per CTA);
// 3 = the fourth operand stays in registers (one ciphertext: 8 tile pairs per CTA, 224 KiB)
template <int PAIRS, int GROUPS, int NBUF> struct FusedCfg {
    static constexpr int kWarps = PAIRS * GROUPS;
    static constexpr size_t kBufBytes = (size_t)kWarps * NBUF * 512 * sizeof(u64);
    static constexpr size_t kTwBytes = (size_t)PAIRS * 2 * 512 * sizeof(Twiddle);
    static constexpr size_t kSmem = kBufBytes + kTwBytes + (PAIRS + 2 * kWarps) * 16;
};

// MODE 0: tensor product, MODE 1: key-switch inner product.
// The operands' tile pairs (4 KiB contiguous each) are fetched by bulk copies (TMA) straight into the warp's buffers: one exposed
// load latency per ciphertext instead of one per operand, and no registers held across it.
template <int KA, int HB, bool NEAR, int PAIRS, int GROUPS, int NBUF, int MODE>
__global__ void __launch_bounds__(32 * PAIRS * GROUPS, 1) bal_fused_kernel(const FusedArgs a) {
    using B = BalB<HB, NEAR>;
    using Cfg = FusedCfg<PAIRS, GROUPS, NBUF>;
    extern __shared__ __align__(128) unsigned char raw[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t pl_local = warp % PAIRS, gl = warp / PAIRS;
    u64* buf = reinterpret_cast<u64*>(raw) + (size_t)warp * NBUF * 512;
    Twiddle* sbf = reinterpret_cast<Twiddle*>(raw + Cfg::kBufBytes) + (size_t)pl_local * 1024;
    Twiddle* sbi = sbf + 512;
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw + Cfg::kBufBytes + Cfg::kTwBytes);
    uint64_t* bar = bars + pl_local * 2;                               // twiddle blocks of the pair
    uint64_t* wbar0 = bars + 2 * PAIRS + warp * 4;                     // this warp's operand copies
    uint64_t* wbar1 = wbar0 + 2;
    constexpr uint32_t pairs = 1u << (KA - 1), pair_blocks = pairs / PAIRS;
    const uint32_t gblocks = (a.groups + GROUPS - 1) / GROUPS;
    const uint32_t pb = blockIdx.x % pair_blocks, r = blockIdx.x / pair_blocks, gb = r % gblocks, limb = r / gblocks;
    const uint32_t pair = pb * PAIRS + pl_local, grp = gb * GROUPS + gl;
    const uint32_t pl = a.limb_begin + limb;
    if (threadIdx.x < PAIRS) mbar_init(bars + threadIdx.x * 2, 1);
    if (lane == 0) { mbar_init(wbar0, 1); mbar_init(wbar1, 1); }
    __syncthreads();
    if (threadIdx.x == 0) mbar_fence_init();
    __syncthreads();
    if (gl == 0 && lane == 0) {
        mbar_arrive_expect_tx(bar, 2 * 512 * sizeof(Twiddle));
        bulk_copy_g2s(sbf, a.blocks_f + ((size_t)pl * pairs + pair) * 512, 512 * sizeof(Twiddle), bar);
        bulk_copy_g2s(sbi, a.blocks_i + ((size_t)pl * pairs + pair) * 512, 512 * sizeof(Twiddle), bar);
    }
    if (grp >= a.groups) return;
    const LimbParams P = a.params[pl];
    const size_t limb_off = (size_t)limb * a.n + (size_t)pair * 512;
    constexpr int B0 = BalA<KA, HB, NEAR>::fwd_out_bound();
    constexpr uint32_t kTile = 512 * sizeof(u64);
    u64* b0 = buf; u64* b1 = buf + 512; u64* b2 = buf + 1024;
    uint32_t par0 = 0, par1 = 0;
    bool first = true;
#pragma unroll 1
    for (uint32_t ct = grp; ct < a.nb; ct += a.groups) {
        const uint64_t* in = a.in + (size_t)ct * a.in_poly + limb_off;
        uint64_t* out = a.out + (size_t)ct * a.out_poly + limb_off;
        if (MODE == 0) {
            const bool sq = a.square != 0;
            const int staged = sq ? 2 : (NBUF == 4 ? 4 : 2);          // operands fetched now; with three buffers b0 follows b1 through b2
            if (lane == 0) {
                fence_proxy_async_smem();
                mbar_arrive_expect_tx(wbar0, staged * kTile);
                for (int p = 0; p < staged; p++) bulk_copy_g2s(buf + p * 512, in + a.in_plane[p], kTile, wbar0);
            }
            u64 y[16];                                              // three buffers: b1 (a1 when squaring) stays in registers
            if (NBUF == 3 && !sq) {
                u64 x[16];
                B::fwd_load(lane, in + a.in_plane[3], x);
                if (first) mbar_wait(bar, 0);
                B::template fwd_phase1<B0>(lane, x, b2, sbf, P);
                __syncwarp();
                B::template fwd_phase2_regs<B0>(lane, b2, sbf, P, y);
                __syncwarp();
                if (lane == 0) {
                    fence_proxy_async_smem();
                    mbar_arrive_expect_tx(wbar1, kTile);
                    bulk_copy_g2s(b2, in + a.in_plane[2], kTile, wbar1);
                }
            } else if (first) mbar_wait(bar, 0);
            mbar_wait(wbar0, par0); par0 ^= 1;
            B::template fwd_phase1_inplace<B0>(lane, b0, sbf, P);
            B::template fwd_phase1_inplace<B0>(lane, b1, sbf, P);
            if (!sq) {
                if (NBUF == 3) { mbar_wait(wbar1, par1); par1 ^= 1; }
                B::template fwd_phase1_inplace<B0>(lane, b2, sbf, P);
                if (NBUF == 4) B::template fwd_phase1_inplace<B0>(lane, buf + 3 * 512, sbf, P);
            }
            __syncwarp();
            B::template fwd_phase2<B0>(lane, b0, sbf, P);
            if (!sq) {
                B::template fwd_phase2<B0>(lane, b1, sbf, P);
                B::template fwd_phase2<B0>(lane, b2, sbf, P);
                if (NBUF == 4) B::template fwd_phase2<B0>(lane, buf + 3 * 512, sbf, P);
            } else {
                B::template fwd_phase2_regs<B0>(lane, b1, sbf, P, y);
            }
            // the lane's own rows: products in place of the operands
#pragma unroll
            for (int k = 0; k < 8; k++) {
                u64 *p0 = B::row_chunk(lane, b0, k), *p1 = B::row_chunk(lane, b1, k), *p2 = B::row_chunk(lane, b2, k);
                u64 a0x, a0y, a1x, a1y, c0x, c0y, c1x, c1y;
                ld2(p0, a0x, a0y);
                if (!sq) {
                    ld2(p1, a1x, a1y); ld2(p2, c0x, c0y);
                    if (NBUF == 4) ld2(B::row_chunk(lane, buf + 3 * 512, k), c1x, c1y); else { c1x = y[2 * k]; c1y = y[2 * k + 1]; }
                } else { a1x = y[2 * k]; a1y = y[2 * k + 1]; c0x = a0x; c0y = a0y; c1x = a1x; c1y = a1y; }
                u64 hi, lo;
                const u64 r0x = mul_mod(a0x, c0x, P), r0y = mul_mod(a0y, c0y, P);
                const u64 r2x = mul_mod(a1x, c1x, P), r2y = mul_mod(a1y, c1y, P);
                hi = 0; lo = 0; mac128(hi, lo, a0x, c1x); mac128(hi, lo, a1x, c0x); const u64 r1x = barrett128(hi, lo, P.q, P.mu_hi, P.mu_lo);
                hi = 0; lo = 0; mac128(hi, lo, a0y, c1y); mac128(hi, lo, a1y, c0y); const u64 r1y = barrett128(hi, lo, P.q, P.mu_hi, P.mu_lo);
                st2(p0, r0x, r0y); st2(p1, r1x, r1y); st2(p2, r2x, r2y);
            }
        } else {
            const uint64_t* kp = a.key + limb_off + (size_t)lane * 16;
            if (lane == 0) {
                fence_proxy_async_smem();
                mbar_arrive_expect_tx(wbar0, a.dnum * kTile);
                for (uint32_t d = 0; d < a.dnum; d++) bulk_copy_g2s(buf + d * 512, in + a.in_plane[d], kTile, wbar0);
            }
            for (uint32_t d = 0; d < 2 * a.dnum; d++) prefetch_l2(kp + (size_t)d * a.key_poly);      // the lane's 128 bytes of every key polynomial
            if (first) mbar_wait(bar, 0);
            mbar_wait(wbar0, par0); par0 ^= 1;
            for (uint32_t d = 0; d < a.dnum; d++) B::template fwd_phase1_inplace<B0>(lane, buf + d * 512, sbf, P);
            __syncwarp();
            for (uint32_t d = 0; d < a.dnum; d++) B::template fwd_phase2<B0>(lane, buf + d * 512, sbf, P);
#pragma unroll 2
            for (int k = 0; k < 8; k++) {
                u64 h0x = 0, l0x = 0, h0y = 0, l0y = 0, h1x = 0, l1x = 0, h1y = 0, l1y = 0;
                for (uint32_t d = 0; d < a.dnum; d++) {
                    u64 vx, vy, kbx, kby, kax, kay;
                    ld2(B::row_chunk(lane, buf + d * 512, k), vx, vy);
                    ldg2(kp + (size_t)(2 * d) * a.key_poly + 2 * k, kbx, kby);
                    ldg2(kp + (size_t)(2 * d + 1) * a.key_poly + 2 * k, kax, kay);
                    mac128(h0x, l0x, vx, kbx); mac128(h0y, l0y, vy, kby);
                    mac128(h1x, l1x, vx, kax); mac128(h1y, l1y, vy, kay);
                }
                st2(B::row_chunk(lane, b0, k), barrett128(h0x, l0x, P.q, P.mu_hi, P.mu_lo), barrett128(h0y, l0y, P.q, P.mu_hi, P.mu_lo));
                st2(B::row_chunk(lane, b1, k), barrett128(h1x, l1x, P.q, P.mu_hi, P.mu_lo), barrett128(h1y, l1y, P.q, P.mu_hi, P.mu_lo));
            }
        }
        first = false;
        // inverse tile pass of the results (2 or 3 polynomials), straight from the rows just written
        constexpr int NOUT = MODE == 0 ? 3 : 2;
#pragma unroll
        for (int q = 0; q < NOUT; q++) B::inv_phase2(lane, buf + q * 512, sbi, P);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NOUT; q++) B::inv_phase3(lane, out + a.out_plane[q], buf + q * 512, sbi, P);
        __syncwarp();
    }
}

template <int KA, int HB, bool NEAR, int MODE>
static int run_fused(fhe_b200_plan* plan, FusedArgs a, cudaStream_t st) {
    static PerDeviceOnce attr_once;
    constexpr int NB2 = MODE == 0 ? 4 : 3;           // buffers per warp of the two-group configuration
    if (attr_once.need(plan->device)) {
        FHE_CUDA(cudaFuncSetAttribute(bal_fused_kernel<KA, HB, NEAR, 4, 2, NB2, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FusedCfg<4, 2, NB2>::kSmem));
        FHE_CUDA(cudaFuncSetAttribute(bal_fused_kernel<KA, HB, NEAR, 8, 1, 3, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FusedCfg<8, 1, 3>::kSmem));
    }
    constexpr uint32_t pairs = 1u << (KA - 1);
    const bool prof = profile_on();
    if (prof) profile_begin(MODE == 0 ? 5 : 6, (uint64_t)a.limb_count * a.nb, st);
    if (a.nb >= 2) {
        // two warps per tile pair share its twiddle blocks; more groups only while the grid would otherwise not fill the SMs
        uint32_t groups = 2;
        const uint32_t ctas = a.limb_count * (pairs / 4);
        while (groups + 2 <= a.nb && (uint64_t)ctas * (groups / 2) < (uint64_t)2 * plan->sm_count) groups += 2;
        a.groups = groups;
        const uint32_t grid = a.limb_count * (pairs / 4) * ((groups + 1) / 2);
        bal_fused_kernel<KA, HB, NEAR, 4, 2, NB2, MODE><<<grid, 256, FusedCfg<4, 2, NB2>::kSmem, st>>>(a);
    } else {
        a.groups = 1;
        const uint32_t grid = a.limb_count * (pairs / 8);
        bal_fused_kernel<KA, HB, NEAR, 8, 1, 3, MODE><<<grid, 256, FusedCfg<8, 1, 3>::kSmem, st>>>(a);
    }
    if (prof) profile_end(st);
    FHE_LAUNCH_CHECK();
    return 0;
}

template <int HB, bool NEAR, int MODE>
static int dispatch_fused(fhe_b200_plan* plan, const FusedArgs& a, cudaStream_t st) {
    switch (plan->logn) {
        case 13: return run_fused<5, HB, NEAR, MODE>(plan, a, st);
        case 14: return run_fused<6, HB, NEAR, MODE>(plan, a, st);
        case 15: return run_fused<7, HB, NEAR, MODE>(plan, a, st);
        case 16: return run_fused<8, HB, NEAR, MODE>(plan, a, st);
    }
    set_error("fused tile pass: unsupported ring degree 2^%u", plan->logn);
    return FHE_B200_EINVAL;
}

bool fused_tile_supported(const fhe_b200_plan* plan, uint32_t dnum) { return plan->bal && dnum <= 3; }

// mode 0: tensor product (4 operands, or 2 when square), mode 1: inner product of dnum digits with the key
int launch_fused_tile(fhe_b200_plan* plan, int mode, const FusedTile& t, cudaStream_t st) {
    FHE_TRY(check_range(plan, t.nb, t.limb_begin, t.limb_count));
    FHE_REQUIRE(plan->bal, "fused tile pass: needs the balanced two-pass NTT (2^13 <= N <= 2^16)");
    FHE_REQUIRE(mode == 0 || (t.dnum >= 1 && t.dnum <= 3 && t.key), "fused tile pass: at most 3 key-switch digits");
    if (!t.nb || !t.limb_count) return 0;
    DeviceGuard dev_guard(plan->device);
    FusedArgs a;
    a.out = t.out; a.in = t.in;
    a.blocks_f = plan->d_fwd_bal; a.blocks_i = plan->d_inv_bal; a.params = plan->d_params;
    a.n = plan->n; a.limb_begin = t.limb_begin; a.limb_count = t.limb_count; a.nb = t.nb; a.groups = 1;
    for (int i = 0; i < 4; i++) a.in_plane[i] = t.in_plane[i];
    for (int i = 0; i < 3; i++) a.out_plane[i] = t.out_plane[i];
    a.in_poly = t.in_poly; a.out_poly = t.out_poly;
    a.key = t.key; a.key_poly = t.key_poly; a.dnum = t.dnum; a.square = t.square ? 1 : 0;
    if (mode == 0)
        return plan->near60 ? dispatch_fused<16, true, 0>(plan, a, st)
             : plan->hb == 16 ? dispatch_fused<16, false, 0>(plan, a, st) : dispatch_fused<8, false, 0>(plan, a, st);
    return plan->near60 ? dispatch_fused<16, true, 1>(plan, a, st)
         : plan->hb == 16 ? dispatch_fused<16, false, 1>(plan, a, st) : dispatch_fused<8, false, 1>(plan, a, st);
}

}  // namespace fhe_b200
