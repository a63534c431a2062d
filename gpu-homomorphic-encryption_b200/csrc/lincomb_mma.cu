// RNS linear combination with the matrix product on the tensor cores (int8-decomposed, mma.sync m16n8k32 u8 x u8 -> s32).
//
// Same primitive, same constants and the same bits as lincomb_kernel (lincomb.cu; replaces fast_base_conversion_kernel,
// /root/reference/include/rns.cuh:116-125): only  sum_i z_i * M[i][k]  is computed differently.  Every z_i and every matrix
// entry is cut into its 8 bytes; the product of byte a of z_i and byte b of M[i][k] has weight 2^(8(a+b)), so with the matrix
// laid out Toeplitz-fashion -- column (k, c) of operand B holds byte (c - a) of M[i][k] in row (i, a) -- one integer GEMM
//     C[coefficient][(k, c)] = sum_{(i,a)} A[coefficient][(i, a)] * B[(i, a)][(k, c)],   c = 0..14,
// yields the 15 partial sums of every 128-bit dot product, each below S * 8 * 255^2 < 2^31.  sum_c C[.][(k,c)] 2^(8c) is the
// exact 128-bit value the IMAD kernel accumulates; Barrett reduction and the epilogues are shared code.  (Half of B is zero:
// 128 byte-MACs per 64-bit MAC.  The tensor pipe still wins over four IMAD.WIDE per MAC -- DESIGN.md section 4.)
//
// Work split: a warp owns 16 consecutive coefficients (the M dimension of the mma).  Lane (g, q) = (lane / 4, lane % 4) loads
// sources 4*kt + q of rows g and g + 8: the two 32-bit halves of a 64-bit residue ARE the A fragments (k is ordered
// [source q][byte 0..3] then [source q][byte 4..7] inside a k-tile), so no byte shuffling is needed.  Targets are processed four
// at a time: n-tile t of a group carries columns (k = 4G + q, c = 2t + u), which is exactly what lane (g, q) receives in its
// accumulator fragment -- after 8 n-tiles a lane holds all partial sums of target 4G + q for its two rows and finishes them
// alone (no shuffles).  B fragments are precomputed per (group, k-tile, n-tile, lane) and read through the read-only cache.
#include "lincomb.cuh"
#include "host_math.hpp"

namespace fhe_b200 {

struct LcMmaArgs {
    const u64 *src_mod, *pre, *pre_s, *th_hi, *th_lo;      // [>= S]
    const u64 *dst_mod, *mu_hi, *mu_lo, *c, *lam;          // [T]
    const uint2* bfrag;                                    // [NG][KT][4][32][2]: n-tiles (2j, 2j+1) of a lane adjacent
    LcView v;
    uint32_t S, T, NG, logn, use_pre, use_extra, c_is_one;
    size_t tiles;                                          // batch * n / 16
};

__device__ __forceinline__ void mma_u8(int (&d)[4], const u32 (&a)[4], const uint2 b) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

__device__ __forceinline__ u64 shfl_x64(u64 v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// sum_c p[c] * 2^(8c), c = 0..14, every p[c] < 2^31, as a 128-bit value (the caller guarantees it fits).  Grouped by c mod 4 the
// partial sums are the 32-bit digits of four 128-bit numbers E_r; E_r << 8r is four funnel shifts, and the three additions are
// plain carry chains -- all on 32-bit registers (written with 64-bit integers the compiler pairs the accumulator registers up
// and spends ~30 moves per value doing so).
__device__ __forceinline__ void assemble128(const u32 (&p)[16], u64& hi, u64& lo) {
    u32 v0 = p[0], v1 = p[4], v2 = p[8], v3 = p[12];
#pragma unroll
    for (int r = 1; r < 4; r++) {
        const u32 d0 = p[r], d1 = p[r + 4], d2 = p[r + 8], d3 = (r + 12 < 15) ? p[r + 12] : 0u;
        const u32 w0 = d0 << (8 * r), w1 = __funnelshift_l(d0, d1, 8 * r), w2 = __funnelshift_l(d1, d2, 8 * r),
                  w3 = __funnelshift_l(d2, d3, 8 * r);
        asm("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %6;\n\taddc.u32 %3, %3, %7;"
            : "+r"(v0), "+r"(v1), "+r"(v2), "+r"(v3) : "r"(w0), "r"(w1), "r"(w2), "r"(w3));
    }
    lo = ((u64)v1 << 32) | v0; hi = ((u64)v3 << 32) | v2;
}

// Persistent CTAs: the whole B-fragment table (NG * KT * 2 KiB) is staged in shared memory once per CTA, then the CTA's warps
// walk the coefficient tiles with a grid stride.  (Reading the fragments through L1 instead left the tensor pipe 29% busy with
// long_scoreboard the top stall: ncu, profiles/r01_lincomb_mma_ncu_summary.md.)
template <int KT>
__global__ void __launch_bounds__(512) lincomb_mma_kernel(const LcMmaArgs a) {
    extern __shared__ __align__(16) unsigned char lc_smem[];
    uint2* sB = reinterpret_cast<uint2*>(lc_smem);
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.bfrag);
        uint4* dst = reinterpret_cast<uint4*>(lc_smem);
        const uint32_t n16 = a.NG * KT * 8 * 32 / 2;
        for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    // per-source constants [5][4 KT] (modulus, pre, pre', theta hi/lo), per-source limb offsets [4 KT], per-target constants after
    // the table: no dependent global loads inside the tile loop
    constexpr int SP = 4 * KT;
    u64* sSrc = reinterpret_cast<u64*>(lc_smem + (size_t)a.NG * KT * 8 * 32 * sizeof(uint2));
    u64* sOff = sSrc + 5 * SP;                          // element offset of source limb i inside a polynomial
    u64* sCpy = sOff + SP;                              // [2][SP]: address of source limb i in the pass-through output (polynomial 0), words per polynomial
    u64* sDst = sCpy + 2 * SP;                          // [6][T]: modulus, mu_hi, mu_lo, c, lam, epilogue scalar
    u64* sIdx = sDst + 6 * (size_t)a.T;                 // [4][T]: address of the target limb in out (polynomial 0), element offsets in extra / epilogue operands, words per polynomial of out
    const size_t nn = (size_t)1 << a.logn;
    for (uint32_t i = threadIdx.x; i < (uint32_t)SP; i += blockDim.x) {
        const bool in = i < a.S;
        sSrc[i] = in ? a.src_mod[i] : 3; sSrc[SP + i] = in ? a.pre[i] : 0; sSrc[2 * SP + i] = in ? a.pre_s[i] : 0;
        sSrc[3 * SP + i] = in ? a.th_hi[i] : 0; sSrc[4 * SP + i] = in ? a.th_lo[i] : 0;
        sOff[i] = in ? (u64)a.v.src_idx[i] * nn : 0;
        if (in && a.v.copy_out) {
            if (a.v.copy_tab) { sCpy[SP + i] = a.v.copy_tab[2 * i + 1]; sCpy[i] = a.v.copy_tab[2 * i] + a.v.copy_poly0 * sCpy[SP + i] * 8; }
            else { sCpy[SP + i] = a.v.copy_stride; sCpy[i] = (u64)(a.v.copy_out + (size_t)a.v.copy_idx[i] * nn); }
        } else { sCpy[i] = 0; sCpy[SP + i] = 0; }
    }
    for (uint32_t k = threadIdx.x; k < a.T; k += blockDim.x) {
        sDst[k] = a.dst_mod[k]; sDst[a.T + k] = a.mu_hi[k]; sDst[2 * a.T + k] = a.mu_lo[k]; sDst[3 * a.T + k] = a.c[k];
        sDst[4 * a.T + k] = a.lam[k]; sDst[5 * a.T + k] = a.v.epi_scalar ? a.v.epi_scalar[k] : 0;
        if (a.v.out_tab) { sIdx[3 * a.T + k] = a.v.out_tab[2 * k + 1]; sIdx[k] = a.v.out_tab[2 * k] + a.v.out_poly0 * sIdx[3 * a.T + k] * 8; }
        else { sIdx[3 * a.T + k] = a.v.out_stride; sIdx[k] = (u64)(a.v.out + (size_t)a.v.dst_idx[k] * nn); }
        sIdx[a.T + k] = (u64)a.v.extra_idx[k] * nn; sIdx[2 * a.T + k] = (u64)a.v.epi_idx[k] * nn;
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const size_t warps_total = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t tile = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < a.tiles; tile += warps_total) {
    const size_t coef0 = tile * 16;
    const uint32_t b = (uint32_t)(coef0 >> a.logn);
    const uint32_t j0 = (uint32_t)(coef0 & (nn - 1)) + g;            // rows g and g + 8 of the tile
    const u64* inb = a.v.in + (size_t)b * a.v.in_stride;

    // ---- prologue: z = x * pre mod s, fixed-point sums of z * theta (rows 0/1 of this lane, sources 4*kt + q)
    u32 A[KT][4];
    u64 f0[2] = {0, 0}, f1[2] = {0, 0}, f2[2] = {0, 0};
    u64 x[KT][2];
#pragma unroll
    for (int kt = 0; kt < KT; kt++) {                                // all loads first (padding sources read limb 0 and are zeroed)
        const u64* p = inb + sOff[4 * kt + q] + j0;
        x[kt][0] = p[0]; x[kt][1] = p[8];
    }
#pragma unroll
    for (int kt = 0; kt < KT; kt++) {
        const uint32_t i = 4 * kt + q;
        if (i >= a.S) { x[kt][0] = 0; x[kt][1] = 0; }
        else if (a.v.copy_out) {
            u64* o = reinterpret_cast<u64*>(sCpy[i]) + (size_t)b * sCpy[SP + i] + j0;
            o[0] = x[kt][0]; o[8] = x[kt][1];
        }
        if (a.use_pre) {
            const u64 pr = sSrc[SP + i], prs = sSrc[2 * SP + i], sm = sSrc[i];
            x[kt][0] = shoup_mul(x[kt][0], pr, prs, sm); x[kt][1] = shoup_mul(x[kt][1], pr, prs, sm);
        }
        const u64 th = sSrc[3 * SP + i], tl = sSrc[4 * SP + i];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            u64 ph, pl;
            mul128(x[kt][r], th, ph, pl);
            add192(f2[r], f1[r], f0[r], ph, pl);
            add192(f2[r], f1[r], f0[r], 0, mulhi64(x[kt][r], tl));
        }
        A[kt][0] = (u32)x[kt][0]; A[kt][1] = (u32)x[kt][1]; A[kt][2] = (u32)(x[kt][0] >> 32); A[kt][3] = (u32)(x[kt][1] >> 32);
    }
    u64 I_hi[2], I_lo[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
#pragma unroll
        for (int m = 1; m <= 2; m <<= 1) {                           // the four lanes of a row hold disjoint sources
            const u64 o0 = shfl_x64(f0[r], m), o1 = shfl_x64(f1[r], m), o2 = shfl_x64(f2[r], m);
            add192(f2[r], f1[r], f0[r], o1, o0); f2[r] += o2;
        }
        add192(f2[r], f1[r], f0[r], 0, 1ull << 63);
        I_hi[r] = f2[r]; I_lo[r] = f1[r];
    }

    // ---- four targets at a time
    for (uint32_t G = 0; G < a.NG; G++) {
        __syncwarp();                                                // mma.sync needs the whole warp (a ragged last group diverges below)
        int acc[8][4];
#pragma unroll
        for (int t = 0; t < 8; t++) { acc[t][0] = 0; acc[t][1] = 0; acc[t][2] = 0; acc[t][3] = 0; }
        // epilogue operands of this lane's target (extra limb, ModDown minuend and addend): issued now, consumed after the mma block
        const uint32_t k = 4 * G + q;
        const bool live = k < a.T;
        u64 ex[2] = {0, 0}, su[2] = {0, 0}, ad[2] = {0, 0};
        if (live) {
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const uint32_t j = j0 + 8 * r;
                if (a.use_extra) ex[r] = a.v.extra[(size_t)b * a.v.extra_stride + sIdx[a.T + k] + j];
                if (a.v.sub) {
                    const size_t eo = sIdx[2 * a.T + k] + j;
                    su[r] = a.v.sub[(size_t)b * a.v.sub_stride + eo];
                    if (a.v.add) ad[r] = a.v.add[(size_t)b * a.v.add_stride + eo];
                }
            }
        }
        const uint4* bf = reinterpret_cast<const uint4*>(sB) + (size_t)G * KT * 4 * 32 + lane;       // two n-tiles per 16-byte load
#pragma unroll
        for (int kt = 0; kt < KT; kt++) {
#pragma unroll
            for (int t2 = 0; t2 < 4; t2++) {
                const uint4 bb = bf[(kt * 4 + t2) * 32];
                mma_u8(acc[2 * t2], A[kt], make_uint2(bb.x, bb.y));
                mma_u8(acc[2 * t2 + 1], A[kt], make_uint2(bb.z, bb.w));
            }
        }
        if (!live) continue;
        const u64 m = sDst[k], mh = sDst[a.T + k], ml = sDst[2 * a.T + k];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            // partial sum of weight 2^(8c) is acc[c / 2][2 r + (c & 1)], c = 0..14
            u32 pc[16];
#pragma unroll
            for (int cc = 0; cc < 16; cc++) pc[cc] = (u32)acc[cc >> 1][2 * r + (cc & 1)];
            u64 ah, al;
            assemble128(pc, ah, al);
            const uint32_t j = j0 + 8 * r;
            if (a.c_is_one) add128(ah, al, I_hi[r], I_lo[r]); else mac128(ah, al, I_lo[r], sDst[3 * a.T + k]);
            if (a.use_extra) mac128(ah, al, ex[r], sDst[4 * a.T + k]);
            u64 res = barrett128(ah, al, m, mh, ml);
            if (a.v.sub) {
                const u64 d = sub_mod(su[r], res, m);
                u64 ph, pl;
                mul128(d, sDst[5 * a.T + k], ph, pl);
                res = barrett128(ph, pl, m, mh, ml);
                if (a.v.add) res = add_mod(res, ad[r], m);
            }
            reinterpret_cast<u64*>(sIdx[k])[(size_t)b * sIdx[3 * a.T + k] + j] = res;
        }
    }
    __syncwarp();
  }
}

static const int kMmaKT[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 13, 16};

uint32_t lincomb_mma_pad_kt(uint32_t S) {
    const uint32_t need = (S + 3) / 4;
    for (int kt : kMmaKT) if ((uint32_t)kt >= need) return (uint32_t)kt;
    return 0;
}

// host: B fragments [NG][KT][4 n-tile pairs][32 lanes][2] of the Toeplitz byte matrix (see the header comment)
void lincomb_mma_build_bfrag(const LincombConsts& h, uint32_t KT, std::vector<uint2>& out) {
    const uint32_t NG = (h.T + 3) / 4;
    out.assign((size_t)NG * KT * 8 * 32, make_uint2(0, 0));
    auto mbyte = [&](uint32_t i, uint32_t k, int bi) -> uint32_t {
        if (i >= h.S || k >= h.T || bi < 0 || bi > 7) return 0;
        return (uint32_t)((h.M[(size_t)i * h.T + k] >> (8 * bi)) & 0xff);
    };
    for (uint32_t G = 0; G < NG; G++)
        for (uint32_t kt = 0; kt < KT; kt++)
            for (uint32_t t = 0; t < 8; t++)
                for (uint32_t lane = 0; lane < 32; lane++) {
                    const uint32_t gq = lane & 3, col = lane >> 2;            // B fragment: k rows 4*gq.. (+16), column n = lane / 4
                    const uint32_t i = 4 * kt + gq;                           // source
                    const uint32_t k = 4 * G + (col >> 1);                    // target of column n = 2 q + u
                    const int c = (int)(2 * t + (col & 1));                   // weight index
                    uint32_t b0 = 0, b1 = 0;
                    for (int by = 0; by < 4; by++) {
                        b0 |= mbyte(i, k, c - by) << (8 * by);                // rows (i, a = by)
                        b1 |= mbyte(i, k, c - 4 - by) << (8 * by);            // rows (i, a = 4 + by)
                    }
                    out[((((size_t)G * KT + kt) * 4 + t / 2) * 32 + lane) * 2 + (t & 1)] = make_uint2(b0, b1);   // n-tiles 2j, 2j+1 adjacent
                }
}

template <int KT>
static int launch_kt(const LcMmaArgs& a, int sm_count, int device, cudaStream_t st) {
    const size_t smem = (size_t)a.NG * KT * 8 * 32 * sizeof(uint2) + (size_t)(8 * 4 * KT + 10 * a.T) * sizeof(u64);
    FHE_REQUIRE(smem <= 220 * 1024, "lincomb (tensor-core path): fragment table of %zu bytes does not fit shared memory", smem);
    static size_t attr_smem[64] = {0};                   // per device ordinal: function attributes are per device
    if (smem > attr_smem[device & 63]) {
        FHE_CUDA(cudaFuncSetAttribute(lincomb_mma_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem[device & 63] = smem;
    }
    // as many resident warps as the table and the registers allow: 256-thread CTAs (two per SM at 128 registers, three for the
    // k-tile counts that compile to 80), or one CTA of 512 threads when the table needs more than half of the shared memory
    const bool one = smem > 100 * 1024;
    const uint32_t threads = one ? 512 : 256;
    const size_t want = (a.tiles * 32 + threads - 1) / threads;
    int per_sm = 1;                                      // resident 256-thread CTAs per SM (3 for the small tables of ModUp / ModDown)
    if (!one) {
        FHE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lincomb_mma_kernel<KT>, 256, smem));
        if (per_sm < 1) per_sm = 1;
        const char* ev = getenv("FHE_B200_LINCOMB_CTAS");             // experiments: cap the resident CTAs per SM
        if (ev && atoi(ev) > 0 && atoi(ev) < per_sm) per_sm = atoi(ev);
    }
    const size_t cap = (size_t)sm_count * (one ? 1 : per_sm);
    lincomb_mma_kernel<KT><<<(unsigned)(want < cap ? want : cap), threads, smem, st>>>(a);
    FHE_LAUNCH_CHECK();
    return 0;
}

int lincomb_mma_launch(fhe_b200_lincomb* lc, const LcView& view, uint32_t n, uint32_t batch, cudaStream_t st) {
    LcMmaArgs a;
    a.src_mod = lc->src_mod; a.pre = lc->pre; a.pre_s = lc->pre_s; a.th_hi = lc->th_hi; a.th_lo = lc->th_lo;
    a.dst_mod = lc->dst_mod; a.mu_hi = lc->mu_hi; a.mu_lo = lc->mu_lo; a.c = lc->c; a.lam = lc->lam;
    a.bfrag = lc->d_bfrag;
    a.v = view;
    a.S = lc->S; a.T = lc->T; a.NG = (lc->T + 3) / 4; a.logn = host::ilog2(n);
    a.use_pre = lc->use_pre; a.use_extra = lc->use_extra; a.c_is_one = lc->use_pre ? 0 : 1;
    a.tiles = (size_t)batch * n / 16;
    switch (lc->mma_kt) {
#define KT_CASE(N_) case N_: return launch_kt<N_>(a, lc->sm_count, lc->device, st);
        KT_CASE(1) KT_CASE(2) KT_CASE(3) KT_CASE(4) KT_CASE(5) KT_CASE(6) KT_CASE(7) KT_CASE(8) KT_CASE(10) KT_CASE(13) KT_CASE(16)
#undef KT_CASE
    }
    set_error("lincomb (tensor-core path): unsupported k-tile count %u", lc->mma_kt);
    return FHE_B200_EINVAL;
}

}  // namespace fhe_b200
