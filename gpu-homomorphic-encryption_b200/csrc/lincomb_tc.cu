// RNS linear combination with the matrix product on the 5th-generation tensor cores: tcgen05.mma.kind::i8, accumulators in TMEM.
//
// Same primitive, constants and bits as lincomb_kernel / lincomb_mma_kernel (replaces fast_base_conversion_kernel,
// /root/reference/include/rns.cuh:116-125): residues and matrix entries are cut into bytes and the matrix is laid out
// Toeplitz-fashion, so that ONE u8 x u8 -> s32 GEMM yields the partial sums p_c = sum over byte pairs of weight 2^(8c), c = 0..14, of
// every 128-bit dot product (lincomb_mma.cu explains the decomposition).  What changes is who multiplies:
//   lincomb_mma_kernel : mma.sync m16n8k32 -- a warp owns 16 coefficients, issues 300 blocking MMAs per tile and holds the
//                        accumulators in registers; tensor pipe 38-43 % busy, the warp's scalar work waits behind its own MMAs;
//   this kernel        : a group of four warps owns 128 coefficients (one TMEM lane each).  Every thread prepares the bytes of ITS
//                        coefficient (x -> x * pre mod s, fixed-point overflow sum) and stores them as one row of the A operand in
//                        shared memory (canonical K-major core-matrix layout, 16-byte stores, conflict-free); one elected thread
//                        issues the MMAs (M = 128, N = 16 columns per target, 8 targets per chunk, K = 32 bytes per instruction)
//                        against the B operand that the CTA keeps in shared memory for its whole life; the accumulators land in
//                        the group's 128 TMEM columns and tcgen05.commit arrives on an mbarrier; each thread reads the sixteen
//                        partial sums of a target for its coefficient with one tcgen05.ld (32x32b.x16), assembles the 128-bit
//                        value, reduces and stores.  The tensor core works on one group's chunk while the other three groups of
//                        the SM run their scalar work: nothing blocks behind an MMA, and loads / stores are 256 bytes per warp.
// The fixed-point sum that yields the rounded integer I rides on the same GEMM: two pseudo-targets in front of the real ones carry
// theta_hi and theta_lo as their "matrix entries", so the tensor core returns H = sum z_i th_hi_i and G = sum z_i th_lo_i exactly; the
// specification wants sum floor(z_i th_lo_i / 2^64), which is (G - F) / 2^64 with F = sum (z_i th_lo_i mod 2^64) -- and since G = F modulo
// 2^64, all a thread has to compute is the number of carries of that sum of low words: one mul.lo and one add-with-carry per source
// instead of a 128-bit product, a mul.hi and two 192-bit additions.  Same I, bit for bit, as lincomb_kernel / lincomb_mma_kernel.
// Persistent CTAs, one per SM: 4 groups x 128 threads, the whole TMEM (4 x 128 columns).
//
// FOLD form (every target modulus m in (2^60 - 2^32, 2^60): the whole prime chain).  The Toeplitz layout spends 16 columns per target and
// returns a 128-bit sum that then needs a full 128 -> 64 bit reduction per target and coefficient -- and that scalar epilogue, not the
// GEMM, is what the kernel's time goes to.  Modular reduction commutes with the byte split of z:
//   sum_i z_i M_ik = sum_i sum_a z_ia (M_ik 2^(8a))  =  sum_(i,a) z_ia F_(i,a),k   (mod m_k),   F_(i,a),k = M_ik 2^(8a) mod m_k  < 2^60,
// so with K row (i, a) carrying the BYTES of F_(i,a),k in eight columns (no zeros, half the columns, half the tensor-core time) the
// accumulators are p_b = sum z_ia byte_b(F) < 2^25 and V = sum_b p_b 2^(8b) < 2^82: six IMAD.WIDE assemble it, and because
// 2^60 = delta (mod m) one more folds it below 2m:  V = (V mod 2^60) + (V >> 60) delta.  No Montgomery factor, no 128-bit product:
// about 60 issue cycles per target instead of about 105, two chunks of targets per tile instead of four for Q -> R.
#include <type_traits>
#include "lincomb.cuh"
#include "host_math.hpp"
#include "tma.cuh"

namespace fhe_b200 {

constexpr int kTcGroups = 4;             // warp groups per CTA (each owns 128 TMEM columns)
// column plan of a group's 128 TMEM columns: chunk 0 = the two pseudo-targets (2 x 16 columns) + the first targets, later chunks targets only
__host__ __device__ constexpr uint32_t tc_wt(bool fold) { return fold ? 8u : 16u; }                 // columns per target
__host__ __device__ constexpr uint32_t tc_first(bool fold) { return (128u - 32u) / tc_wt(fold); }   // targets in chunk 0
__host__ __device__ constexpr uint32_t tc_later(bool fold) { return 128u / tc_wt(fold); }           // targets in a later chunk
__host__ __device__ inline uint32_t tc_chunks(uint32_t T, bool fold) {
    return T <= tc_first(fold) ? 1u : 1u + (T - tc_first(fold) + tc_later(fold) - 1) / tc_later(fold);
}
__host__ __device__ inline uint32_t tc_pad16(uint32_t c) { return (c + 15u) & ~15u; }
// total (padded) columns of all chunks: the B operand is K rows of that many bytes
__host__ __device__ inline uint32_t tc_total_cols(uint32_t T, bool fold) {
    uint32_t tot = 0, t0 = 0;
    for (uint32_t ch = 0; ch < tc_chunks(T, fold); ch++) {
        const uint32_t cap = ch == 0 ? tc_first(fold) : tc_later(fold), cnt = T - t0 < cap ? T - t0 : cap;
        tot += tc_pad16((ch == 0 ? 32u : 0u) + cnt * tc_wt(fold));
        t0 += cnt;
    }
    return tot;
}

struct LcTcArgs {
    const u64 *src_mod, *pre, *pre_s, *th_hi, *th_lo;      // [>= S]
    const u64 *dst_mod, *mu_hi, *mu_lo, *c, *lam;          // [T]
    const uint8_t* bmat;                                   // B operand, canonical layout per chunk (lincomb_tc_build_b)
    const u64* fold;                                       // FOLD form: [T] Shoup companion of lam
    LcView v;
    uint32_t S, T, KS, logn, use_pre, use_extra, c_is_one; // KS = k-steps of 32 bytes (4 sources each)
    size_t tiles;                                          // batch * n / 128
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void group_barrier(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle: 8-row x 16-byte core matrices, rows 16 bytes apart;
// sbo = bytes between 8-row groups, lbo = bytes between the two 16-byte K chunks of one instruction
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) |
           (1ull << 46);                                   // descriptor version 1 (sm_100), base offset 0, layout type 0 = no swizzle
}
// instruction descriptor of kind::i8: D = s32, A = B = u8, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t tc_idesc(uint32_t n) {
    return (2u << 4) | (0u << 7) | (0u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, u32 (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, u32 (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// sum_c p[c] * 2^(8c), c = 0..14 (p[15] = 0), every p[c] < 2^31, as a 128-bit value: see assemble128 in lincomb_mma.cu
__device__ __forceinline__ void tc_assemble128(const u32 (&p)[16], u64& hi, u64& lo) {
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

// the same sum as 160 bits (five 32-bit words): the pseudo-targets' entries are full 64-bit words, so the sum can pass 2^128
__device__ __forceinline__ void tc_assemble160(const u32 (&p)[16], u64& w2, u64& w1, u64& w0) {
    u32 v0 = p[0], v1 = p[4], v2 = p[8], v3 = p[12], v4 = 0;
#pragma unroll
    for (int r = 1; r < 4; r++) {
        const u32 d0 = p[r], d1 = p[r + 4], d2 = p[r + 8], d3 = (r + 12 < 15) ? p[r + 12] : 0u;
        const u32 x0 = d0 << (8 * r), x1 = __funnelshift_l(d0, d1, 8 * r), x2 = __funnelshift_l(d1, d2, 8 * r),
                  x3 = __funnelshift_l(d2, d3, 8 * r), x4 = d3 >> (32 - 8 * r);
        asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\taddc.cc.u32 %3, %3, %8;\n\taddc.u32 %4, %4, %9;"
            : "+r"(v0), "+r"(v1), "+r"(v2), "+r"(v3), "+r"(v4) : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(x4));
    }
    w0 = ((u64)v1 << 32) | v0; w1 = ((u64)v3 << 32) | v2; w2 = v4;
}

// FOLD form.  The eight columns of a target are stored in the order byte 0, 4, 1, 5, 2, 6, 3, 7, so that the register pair (r0, r1) IS
// p_0 + p_4 2^32 -- the 64-bit addend of the first IMAD.WIDE.  With every p_b < 2^25:
//   lo = p0 + p4 2^32 + p1 2^8 + p2 2^16 + p3 2^24 (+ x0 y0)  < 2^58,      hi = p5 2^8 + p6 2^16 + p7 2^24 (+ x0 y1)  < 2^50,
//   V  = lo + hi 2^32.   (x0, y1:y0) is an optional small product riding on the same chains: the overflow count times c_k.
// The weights 2^8, 2^16, 2^24 come in registers (256 + an opaque zero ...): as literals ptxas rewrites every IMAD.WIDE into
// IMAD.SHL + IMAD.HI + a three-input add, ten issue cycles instead of four.
struct TcW { u32 w8, w16, w24; };
__device__ __forceinline__ TcW tc_weights() {
    const u32 z = (u32)kOpaqueZero;
    return TcW{256u + z, 65536u + z, 16777216u + z};
}
__device__ __forceinline__ void tc_assemble_fold(const u32 (&r)[8], const TcW& w, u32 x0, u64 y, u64& lo, u64& hi) {
    asm("{\n\t"
        ".reg .u64 a, g;\n\t"
        ".reg .u32 y0, y1;\n\t"
        "mov.b64 {y0, y1}, %11;\n\t"
        "mov.b64 a, {%2, %3};\n\t"
        "mad.wide.u32 a, %4, %12, a;\n\t"
        "mad.wide.u32 a, %6, %13, a;\n\t"
        "mad.wide.u32 a, %8, %14, a;\n\t"
        "mad.wide.u32 %0, %10, y0, a;\n\t"
        "mul.wide.u32 g, %5, %12;\n\t"
        "mad.wide.u32 g, %7, %13, g;\n\t"
        "mad.wide.u32 g, %9, %14, g;\n\t"
        "mad.wide.u32 %1, %10, y1, g;\n\t"
        "}" : "=&l"(lo), "=&l"(hi)
        : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(x0), "l"(y), "r"(w.w8), "r"(w.w16), "r"(w.w24));
}
__device__ __forceinline__ void tc_assemble_fold(const u32 (&r)[8], const TcW& w, u64& lo, u64& hi) {
    asm("{\n\t"
        ".reg .u64 a, g;\n\t"
        "mov.b64 a, {%2, %3};\n\t"
        "mad.wide.u32 a, %4, %10, a;\n\t"
        "mad.wide.u32 a, %6, %11, a;\n\t"
        "mad.wide.u32 %0, %8, %12, a;\n\t"
        "mul.wide.u32 g, %5, %10;\n\t"
        "mad.wide.u32 g, %7, %11, g;\n\t"
        "mad.wide.u32 %1, %9, %12, g;\n\t"
        "}" : "=&l"(lo), "=&l"(hi)
        : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(w.w8), "r"(w.w16), "r"(w.w24));
}
// V = lo + hi 2^32 (+ e0 + e1 2^64) as three words, folded at 2^60: (V mod 2^60) + (V >> 60) delta, below 2m for V < 2^88
__device__ __forceinline__ u64 tc_fold60(u64 lo, u64 hi, u64 e0, u32 e1, u32 delta) {
    u32 v0, v1, v2;
    asm("{\n\t"
        ".reg .u32 l0, l1, h0, h1, a0, a1;\n\t"
        "mov.b64 {l0, l1}, %3;\n\t"
        "mov.b64 {h0, h1}, %4;\n\t"
        "mov.b64 {a0, a1}, %5;\n\t"
        "add.cc.u32 %0, l0, a0;\n\t"
        "addc.cc.u32 %1, l1, a1;\n\t"
        "addc.u32 %2, h1, %6;\n\t"
        "add.cc.u32 %1, %1, h0;\n\t"
        "addc.u32 %2, %2, 0;\n\t"
        "}" : "=&r"(v0), "=&r"(v1), "=&r"(v2) : "l"(lo), "l"(hi), "l"(e0), "r"(e1));
    const u32 top = __funnelshift_l(v1, v2, 4);                      // V >> 60
    u64 r;
    asm("{\n\t"
        ".reg .u32 m1;\n\t"
        ".reg .u64 m;\n\t"
        "and.b32 m1, %2, 0x0fffffff;\n\t"
        "mov.b64 m, {%1, m1};\n\t"
        "mad.wide.u32 %0, %3, %4, m;\n\t"
        "}" : "=l"(r) : "r"(v0), "r"(v1), "r"(top), "r"(delta));
    return r;
}
__device__ __forceinline__ u64 tc_fold60(u64 lo, u64 hi, u32 delta) {
    u32 v1, v2;
    asm("{\n\t"
        ".reg .u32 l0, l1, h0, h1;\n\t"
        "mov.b64 {l0, l1}, %2;\n\t"
        "mov.b64 {h0, h1}, %3;\n\t"
        "add.cc.u32 %0, l1, h0;\n\t"
        "addc.u32 %1, h1, 0;\n\t"
        "}" : "=&r"(v1), "=&r"(v2) : "l"(lo), "l"(hi));
    const u32 top = __funnelshift_l(v1, v2, 4);
    u64 r;
    asm("{\n\t"
        ".reg .u32 l0, l1, m1;\n\t"
        ".reg .u64 m;\n\t"
        "mov.b64 {l0, l1}, %1;\n\t"
        "and.b32 m1, %2, 0x0fffffff;\n\t"
        "mov.b64 m, {l0, m1};\n\t"
        "mad.wide.u32 %0, %3, %4, m;\n\t"
        "}" : "=l"(r) : "l"(lo), "r"(v1), "r"(top), "r"(delta));
    return r;
}

// shared memory: B operand | A tiles (one per group) | per-source and per-target constants | barriers, TMEM base
// FOLD: every target modulus is in (2^60 - 2^32, 2^60) (see the header); otherwise the Toeplitz form with the 128-bit Barrett epilogue.
// KIND: 0 plain conversion, 1 scale-and-round (extra limb, whole integer I added), 2 conversion with the ModDown epilogue,
//       3 anything else (flags read at run time)
template <bool FOLD, int KIND>
__global__ void __launch_bounds__(128 * kTcGroups, 1) lincomb_tc_kernel(const LcTcArgs a) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    constexpr uint32_t WT = tc_wt(FOLD);
    const uint32_t tid = threadIdx.x, grp = tid >> 7, gtid = tid & 127, warp = tid >> 5;
    const uint32_t K = a.KS * 32;                                   // bytes of one A row
    const uint32_t NT = tc_chunks(a.T, FOLD);                       // chunks
    const size_t b_bytes = (size_t)K * tc_total_cols(a.T, FOLD);
    unsigned char* sB = tc_smem;
    unsigned char* sA = tc_smem + ((b_bytes + 127) & ~(size_t)127) + (size_t)grp * 128 * K;
    u64* sC = reinterpret_cast<u64*>(tc_smem + ((b_bytes + 127) & ~(size_t)127) + (size_t)kTcGroups * 128 * K);
    const uint32_t SP = a.KS * 4;
    u64* sSrc = sC;                                                 // [SP] records of 8 words: modulus, pre | pre', theta_lo | element offset of the limb inside a
                                                                    // polynomial, 0 | pass-through address (polynomial 0), words per polynomial -- read as 16-byte pairs
    u64* sDst = sSrc + 8 * SP;                                      // [7][T]: modulus, mu_hi (FOLD: Shoup companion of lam), mu_lo, c, lam, epilogue scalar and its Shoup companion
    u64* sIdx = sDst + 7 * (size_t)a.T;                             // [4][T]: out address (polynomial 0), extra / epilogue offsets, out words per polynomial
    uint64_t* bars = reinterpret_cast<uint64_t*>(sIdx + 4 * (size_t)a.T);     // one per group
    uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(bars + kTcGroups);
    const size_t nn = (size_t)1 << a.logn;

    {   // stage B and the constants
        const uint4* src = reinterpret_cast<const uint4*>(a.bmat);
        uint4* dst = reinterpret_cast<uint4*>(sB);
        for (size_t i = tid; i < b_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    for (uint32_t i = tid; i < SP; i += blockDim.x) {
        const bool in = i < a.S;
        u64* rec = sSrc + 8 * (size_t)i;
        rec[0] = in ? a.src_mod[i] : 3; rec[1] = in ? a.pre[i] : 0; rec[2] = in ? a.pre_s[i] : 0; rec[3] = in ? a.th_lo[i] : 0;
        rec[4] = in ? (u64)a.v.src_idx[i] * nn : 0; rec[5] = 0; rec[6] = 0; rec[7] = 0;
        if (in && a.v.copy_out) {
            if (a.v.copy_tab) { rec[7] = a.v.copy_tab[2 * i + 1]; rec[6] = a.v.copy_tab[2 * i] + a.v.copy_poly0 * rec[7] * 8; }
            else { rec[7] = a.v.copy_stride; rec[6] = (u64)(a.v.copy_out + (size_t)a.v.copy_idx[i] * nn); }
        }
    }
    for (uint32_t k = tid; k < a.T; k += blockDim.x) {
        sDst[k] = a.dst_mod[k]; sDst[a.T + k] = FOLD ? a.fold[k] : a.mu_hi[k]; sDst[2 * a.T + k] = a.mu_lo[k];
        sDst[3 * a.T + k] = a.c[k]; sDst[4 * a.T + k] = a.lam[k];
        sDst[5 * a.T + k] = a.v.epi_scalar ? a.v.epi_scalar[k] : 0; sDst[6 * a.T + k] = a.v.epi_scalar_shoup ? a.v.epi_scalar_shoup[k] : 0;
        if (a.v.out_tab) { sIdx[3 * a.T + k] = a.v.out_tab[2 * k + 1]; sIdx[k] = a.v.out_tab[2 * k] + a.v.out_poly0 * sIdx[3 * a.T + k] * 8; }
        else { sIdx[3 * a.T + k] = a.v.out_stride; sIdx[k] = (u64)(a.v.out + (size_t)a.v.dst_idx[k] * nn); }
        sIdx[a.T + k] = (u64)a.v.extra_idx[k] * nn; sIdx[2 * a.T + k] = (u64)a.v.epi_idx[k] * nn;
    }
    if (tid < kTcGroups) mbar_init(bars + tid, 1);
    if (warp == 0) {                                                // the whole TMEM: 512 columns, 128 per group
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_base_p)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) mbar_fence_init();
    fence_proxy_async_smem();                                       // B was written with ordinary stores; the tensor core reads it through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_base_p + grp * 128;               // this group's accumulator columns (lane field 0)
    const uint32_t tmem_rd = tmem_d + ((warp & 3u) * 32u << 16);    // this warp's 32 lanes
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    uint32_t parity = 0;
    const TcW wts = tc_weights();
    const bool use_extra = KIND == 3 ? a.use_extra != 0 : KIND == 1;
    const bool c_is_one = KIND == 3 ? a.c_is_one != 0 : KIND == 1;
    const bool has_sub = KIND == 3 ? a.v.sub != nullptr : KIND == 2;
    const bool do_copy = a.v.copy_out != nullptr;
    constexpr bool PRE = FOLD && (KIND == 1 || KIND == 2);          // epilogue operand prefetched a chunk ahead (opv)

#pragma unroll 1
    for (size_t tile = (size_t)blockIdx.x + (size_t)gridDim.x * grp; tile < a.tiles; tile += (size_t)gridDim.x * kTcGroups) {     // CTAs first, then groups: few tiles spread over many SMs
        const size_t coef = tile * 128 + gtid;
        const uint32_t b = (uint32_t)(coef >> a.logn);
        const uint32_t j = (uint32_t)(coef & (nn - 1));
        const u64* inb = a.v.in + (size_t)b * a.v.in_stride + j;
        // the ModDown operands that are loaded pair by pair (minuend / addend) start their way to L2 now, a whole prologue ahead
        if (has_sub)
            for (uint32_t k = 0; k < a.T; k++) {
                if (!PRE) prefetch_l2(a.v.sub + (size_t)b * a.v.sub_stride + sIdx[2 * a.T + k] + j);
                if (a.v.add) prefetch_l2(a.v.add + (size_t)b * a.v.add_stride + sIdx[2 * a.T + k] + j);
            }
        // FOLD form, scale-and-round and ModDown: the epilogue operand of a whole chunk of targets (extra limb / minuend) is loaded into
        // registers a phase ahead (the prologue's registers are free by then): one pair ahead -- the Toeplitz form's scheme -- left a third
        // of the scale-and-round kernel's cycles waiting for them
        u64 opv[16];
        auto load_extras = [&](uint32_t first, uint32_t count) {
            if (PRE) {
#pragma unroll
                for (int t = 0; t < 16; t++)
                    if ((uint32_t)t < count)
                        opv[t] = KIND == 1 ? a.v.extra[(size_t)b * a.v.extra_stride + sIdx[a.T + first + t] + j]
                                           : a.v.sub[(size_t)b * a.v.sub_stride + sIdx[2 * a.T + first + t] + j];
            }
        };

        // ---- prologue: this coefficient's sources -> bytes of z (row gtid of A), fixed-point sum of z * theta
        u64 facc = 0; u32 fcnt = 0;                                 // sum of the low words of z * theta_lo, and its carries
        // batches of eight sources, the loads of the next batch in flight while this one is multiplied (two register sets)
        auto load8 = [&](u64 (&x)[8], uint32_t i0) {
#pragma unroll
            for (int u = 0; u < 8; u++) x[u] = (i0 + u < a.S) ? __ldcg(inb + sSrc[8 * (i0 + u) + 4]) : 0;
        };
        // FULL: all eight sources of the batch exist (no per-source tests); the flags are uniform
        auto work8 = [&](u64 (&x)[8], uint32_t i0, auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t i = i0 + u;
                if (FULL || i < SP) {
                    const ulonglong2 c01 = *reinterpret_cast<const ulonglong2*>(sSrc + 8 * i);          // modulus, pre
                    const ulonglong2 c23 = *reinterpret_cast<const ulonglong2*>(sSrc + 8 * i + 2);      // pre', theta_lo
                    if (do_copy && (FULL || i < a.S)) {
                        const ulonglong2 c67 = *reinterpret_cast<const ulonglong2*>(sSrc + 8 * i + 6);  // pass-through address, words per polynomial
                        reinterpret_cast<u64*>(c67.x)[(size_t)b * c67.y + j] = x[u];
                    }
                    if (a.use_pre) x[u] = csub_sign(shoup_mul_lazy(x[u], c01.y, c23.x, c01.x), c01.x);
                    const u64 lo = x[u] * c23.y;
                    asm("add.cc.u64 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+l"(facc), "+r"(fcnt) : "l"(lo));
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u += 2)
                if (FULL || i0 + u < SP)                            // sources i, i+1 = K bytes [8i, 8i+16) = K chunk i/2
                    *reinterpret_cast<ulonglong2*>(sA + (size_t)((i0 + u) >> 1) * 2048 + gtid * 16) = make_ulonglong2(x[u], x[u + 1]);
        };
        auto work = [&](u64 (&x)[8], uint32_t i0) {
            if (i0 + 8 <= a.S) work8(x, i0, std::true_type{}); else work8(x, i0, std::false_type{});
        };
        {
            u64 xa[8], xb[8];
            load8(xa, 0);
#pragma unroll 1
            for (uint32_t i0 = 0; i0 < SP; i0 += 16) {
                if (i0 + 8 < SP) load8(xb, i0 + 8);
                work(xa, i0);
                if (i0 + 16 < SP) load8(xa, i0 + 16);
                if (i0 + 8 < SP) work(xb, i0 + 8);
            }
        }
        u64 I_hi = 0, I_lo = 0;                                     // known after the first chunk (the two pseudo-targets)
        load_extras(0, min(tc_first(FOLD), a.T));
        fence_proxy_async_smem();                                   // A row visible to the tensor core
        tc_fence_before();
        group_barrier(1 + grp);

        // ---- chunks of targets: MMA into the group's TMEM columns, then every thread finishes its coefficient
        uint32_t t0 = 0, b_off = 0;                                 // first target of the chunk, byte offset of the chunk's B block
#pragma unroll 1
        for (uint32_t ch = 0; ch < NT; ch++) {
            const uint32_t cap = ch == 0 ? tc_first(FOLD) : tc_later(FOLD);
            const uint32_t cnt = min(cap, a.T - t0);                // targets of this chunk
            const uint32_t col0 = ch == 0 ? 32u : 0u;               // TMEM column of its first target
            const uint32_t ncol = tc_pad16(col0 + cnt * WT);
            if (gtid == 0) {
                tc_fence_after();
                const uint32_t idesc = tc_idesc(ncol);
                const uint32_t b_chunk = sB_addr + b_off;
                for (uint32_t ks = 0; ks < a.KS; ks++) {
                    const uint64_t da = tc_smem_desc(sA_addr + ks * 2 * 2048, 2048, 128);
                    const uint64_t db = tc_smem_desc(b_chunk + ks * 2 * ncol * 16, ncol * 16, 128);
                    tc_mma_i8(tmem_d, da, db, idesc, ks > 0 ? 1u : 0u);
                }
                tc_commit(bars + grp);
            }
            mbar_wait(bars + grp, parity); parity ^= 1;
            tc_fence_after();
            auto operands = [&](uint32_t k, u64& ex, u64& su, u64& ad) {
                ex = 0; su = 0; ad = 0;
                if (!PRE && use_extra) ex = a.v.extra[(size_t)b * a.v.extra_stride + sIdx[a.T + k] + j];
                if (has_sub) {
                    const size_t eo = sIdx[2 * a.T + k] + j;
                    if (!PRE) su = a.v.sub[(size_t)b * a.v.sub_stride + eo];
                    if (a.v.add) ad = a.v.add[(size_t)b * a.v.add_stride + eo];
                }
            };
            // ModDown epilogue on a residue that need not be canonical (below 2m): (su - res) * scalar [+ ad]
            auto moddown = [&](uint32_t k, u64 res, u64 m, u64 su, u64 ad) -> u64 {
                const u64 d = su + 2 * m - res;                      // in (0, 3m): the exact Shoup product takes any 64-bit operand
                u64 r = shoup_mul(d, sDst[5 * a.T + k], sDst[6 * a.T + k], m);
                if (a.v.add) r = add_mod(r, ad, m);
                return r;
            };
            auto store = [&](uint32_t k, u64 res) { reinterpret_cast<u64*>(sIdx[k])[(size_t)b * sIdx[3 * a.T + k] + j] = res; };
            // Toeplitz form: sixteen byte-weighted partial sums -> 128 bits -> Barrett
            auto finish16 = [&](uint32_t k, const u32 (&pc)[16], u64 ex, u64 su, u64 ad) {
                const u64 m = sDst[k], mh = sDst[a.T + k], ml = sDst[2 * a.T + k];
                u64 ah, al;
                tc_assemble128(pc, ah, al);
                if (c_is_one) add128(ah, al, I_hi, I_lo); else mac128(ah, al, I_lo, sDst[3 * a.T + k]);
                if (use_extra) mac128(ah, al, ex, sDst[4 * a.T + k]);
                u64 res = barrett128(ah, al, m, mh, ml);
                if (has_sub) {
                    const u64 d = sub_mod(su, res, m);
                    if (a.v.epi_scalar_shoup) res = shoup_mul(d, sDst[5 * a.T + k], sDst[6 * a.T + k], m);
                    else {
                        u64 ph, pl;
                        mul128(d, sDst[5 * a.T + k], ph, pl);
                        res = barrett128(ph, pl, m, mh, ml);
                    }
                    if (a.v.add) res = add_mod(res, ad, m);
                }
                store(k, res);
            };
            // FOLD form: eight partial sums -> below 2^82 -> one fold at 2^60
            auto finish8 = [&](uint32_t k, const u32 (&pc)[8], u64 ex, u64 su, u64 ad) {
                const u64 m = sDst[k];
                const u32 delta = (u32)(0 - m);                      // 2^60 - m: the low word of 2^64 - m
                u64 lo, hi, r;
                if (c_is_one) {                                      // scale-and-round: the whole integer I (65 bits) and extra * lam join V
                    tc_assemble_fold(pc, wts, lo, hi);
                    u64 e0 = I_lo; u32 e1 = (u32)I_hi;
                    if (use_extra) {
                        const u64 e = shoup_mul_lazy3(ex, sDst[4 * a.T + k], sDst[a.T + k], 0 - m);      // below 3m
                        asm("add.cc.u64 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+l"(e0), "+r"(e1) : "l"(e));
                    }
                    r = tc_fold60(lo, hi, e0, e1, delta);
                } else {                                             // conversion: I < 64 (an overflow count) times c_k rides on the assembly
                    tc_assemble_fold(pc, wts, (u32)I_lo, sDst[3 * a.T + k], lo, hi);
                    if (use_extra) {
                        const u64 e = shoup_mul_lazy3(ex, sDst[4 * a.T + k], sDst[a.T + k], 0 - m);
                        r = tc_fold60(lo, hi, e, 0u, delta);
                    } else r = tc_fold60(lo, hi, delta);
                }
                store(k, has_sub ? moddown(k, r, m, su, ad) : csub_sign(r, m));
            };
            // the epilogue operands (extra limb, ModDown minuend / addend) of a pair are loaded one pair ahead: their latency hides
            // behind the previous pair's reductions
            u64 ex0 = 0, su0 = 0, ad0 = 0, ex1 = 0, su1 = 0, ad1 = 0;
            if (cnt >= 2) { operands(t0, ex0, su0, ad0); operands(t0 + 1, ex1, su1, ad1); }
            if (ch == 0) {
                // columns 0..31: H = sum z theta_hi, G = sum z theta_lo (exact, 160 bits each).  sum floor(z theta_lo / 2^64) = (G >> 64) - carries of the
                // threads' sum of low words (G = that sum modulo 2^64);  f = H + that + 2^63,  I = f >> 64.
                u32 p0[16], p1[16];
                tc_ld16(tmem_rd, p0);
                tc_ld16(tmem_rd + 16, p1);
                tc_wait_ld();
                u64 h2, h1, h0, g2, g1, g0;
                tc_assemble160(p0, h2, h1, h0);
                tc_assemble160(p1, g2, g1, g0);
                u64 e1 = g2, e0 = g1;
                asm("sub.cc.u64 %0, %0, %2;\n\tsubc.u64 %1, %1, 0;" : "+l"(e0), "+l"(e1) : "l"((u64)fcnt));
                add192(h2, h1, h0, e1, e0);
                add192(h2, h1, h0, 0, 1ull << 63);
                I_hi = h2; I_lo = h1;
                (void)g0;
            }
            // two targets per iteration: their reductions are independent chains, which keeps the four warps of a scheduler issuing
            if (PRE) {                                               // unrolled: opv is indexed statically
#pragma unroll
                for (int tl = 0; tl < 16; tl += 2) {
                    const uint32_t k = t0 + tl;
                    if ((uint32_t)tl + 1 < cnt) {
                        const u64 csu0 = su0, cad0 = ad0, csu1 = su1, cad1 = ad1;
                        u32 p0[8], p1[8];
                        tc_ld8(tmem_rd + col0 + tl * 8, p0);
                        tc_ld8(tmem_rd + col0 + tl * 8 + 8, p1);
                        if ((uint32_t)tl + 3 < cnt) { operands(k + 2, ex0, su0, ad0); operands(k + 3, ex1, su1, ad1); }
                        tc_wait_ld();
                        finish8(k, p0, KIND == 1 ? opv[tl] : 0, KIND == 2 ? opv[tl] : csu0, cad0);
                        finish8(k + 1, p1, KIND == 1 ? opv[tl + 1] : 0, KIND == 2 ? opv[tl + 1] : csu1, cad1);
                    } else if ((uint32_t)tl < cnt) {                 // an odd target left over
                        operands(k, ex0, su0, ad0);
                        u32 p0[8];
                        tc_ld8(tmem_rd + col0 + tl * 8, p0);
                        tc_wait_ld();
                        finish8(k, p0, KIND == 1 ? opv[tl] : 0, KIND == 2 ? opv[tl] : su0, ad0);
                    }
                }
                // the next chunk's extra limbs: in flight across the barrier and the tensor core's work
                if (ch + 1 < NT) load_extras(t0 + cnt, min(tc_later(FOLD), a.T - t0 - cnt));
            } else {
                uint32_t tl = 0;
#pragma unroll 1
                for (; tl + 1 < cnt; tl += 2) {
                    const uint32_t k = t0 + tl;
                    const u64 cex0 = ex0, csu0 = su0, cad0 = ad0, cex1 = ex1, csu1 = su1, cad1 = ad1;
                    if (FOLD) {
                        u32 p0[8], p1[8];
                        tc_ld8(tmem_rd + col0 + tl * 8, p0);
                        tc_ld8(tmem_rd + col0 + tl * 8 + 8, p1);
                        if (tl + 3 < cnt) { operands(k + 2, ex0, su0, ad0); operands(k + 3, ex1, su1, ad1); }
                        tc_wait_ld();
                        finish8(k, p0, cex0, csu0, cad0);
                        finish8(k + 1, p1, cex1, csu1, cad1);
                    } else {
                        u32 p0[16], p1[16];
                        tc_ld16(tmem_rd + col0 + tl * 16, p0);
                        tc_ld16(tmem_rd + col0 + tl * 16 + 16, p1);
                        if (tl + 3 < cnt) { operands(k + 2, ex0, su0, ad0); operands(k + 3, ex1, su1, ad1); }
                        tc_wait_ld();
                        finish16(k, p0, cex0, csu0, cad0);
                        finish16(k + 1, p1, cex1, csu1, cad1);
                    }
                }
                if (tl < cnt) {                                      // an odd target left over
                    const uint32_t k = t0 + tl;
                    operands(k, ex0, su0, ad0);
                    if (FOLD) {
                        u32 p0[8];
                        tc_ld8(tmem_rd + col0 + tl * 8, p0);
                        tc_wait_ld();
                        finish8(k, p0, ex0, su0, ad0);
                    } else {
                        u32 p0[16];
                        tc_ld16(tmem_rd + col0 + tl * 16, p0);
                        tc_wait_ld();
                        finish16(k, p0, ex0, su0, ad0);
                    }
                }
            }
            t0 += cnt; b_off += K * ncol;
            tc_fence_before();
            group_barrier(1 + grp);                                 // every thread has read its accumulators: the next chunk may overwrite them
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*tmem_base_p) : "memory");
    }
}

// host: B operand, chunk after chunk (tc_chunks).  Canonical K-major layout without swizzle: byte (n, kb) of a chunk with ncol (padded)
// columns at  (kb / 16) * (ncol * 16) + n * 16 + kb % 16;  K row kb = source kb / 8, byte kb % 8 of its z.
//   columns 0..31 of chunk 0: theta_hi, theta_lo (raw 64-bit words: the overflow sum), Toeplitz: column c carries byte c - a of the entry;
//   Toeplitz form: 16 columns per target, the same with the entry M[i][k];
//   FOLD form: 8 columns per target, byte b of F = M[i][k] * 2^(8a) mod m_k in column 2 * (b & 3) + (b >> 2)  (order 0, 4, 1, 5, 2, 6, 3, 7).
void lincomb_tc_build_b(const LincombConsts& h, uint32_t KS, bool fold, std::vector<uint8_t>& out) {
    const uint32_t K = KS * 32, WT = tc_wt(fold);
    out.assign((size_t)K * tc_total_cols(h.T, fold), 0);
    size_t base = 0;
    uint32_t t0 = 0;
    for (uint32_t ch = 0; ch < tc_chunks(h.T, fold); ch++) {
        const uint32_t cap = ch == 0 ? tc_first(fold) : tc_later(fold), cnt = std::min<uint32_t>(cap, h.T - t0);
        const uint32_t col0 = ch == 0 ? 32u : 0u, ncol = tc_pad16(col0 + cnt * WT);
        auto put = [&](uint32_t n, uint32_t kb, uint8_t v) { out[base + (size_t)(kb / 16) * (ncol * 16) + (size_t)n * 16 + kb % 16] = v; };
        for (uint32_t kb = 0; kb < K; kb++) {
            const uint32_t i = kb / 8, aa = kb % 8;
            if (i >= h.S) continue;
            if (ch == 0)
                for (uint32_t ps = 0; ps < 2; ps++) {
                    const uint64_t e = ps == 0 ? h.th_hi[i] : h.th_lo[i];
                    for (uint32_t c = aa; c < aa + 8 && c < 15; c++) put(ps * 16 + c, kb, (uint8_t)((e >> (8 * (c - aa))) & 0xff));
                }
            for (uint32_t tl = 0; tl < cnt; tl++) {
                const uint32_t k = t0 + tl;
                const uint64_t m = h.dst_mod[k], e = h.M[(size_t)i * h.T + k];
                if (fold) {
                    const uint64_t f = host::mulmod(e % m, (uint64_t)((((unsigned __int128)1) << (8 * aa)) % m), m);
                    for (uint32_t bb = 0; bb < 8; bb++) put(col0 + tl * 8 + 2 * (bb & 3) + (bb >> 2), kb, (uint8_t)((f >> (8 * bb)) & 0xff));
                } else
                    for (uint32_t c = aa; c < aa + 8 && c < 15; c++) put(col0 + tl * 16 + c, kb, (uint8_t)((e >> (8 * (c - aa))) & 0xff));
            }
        }
        base += (size_t)K * ncol;
        t0 += cnt;
    }
}

size_t lincomb_tc_smem_bytes(uint32_t S, uint32_t T, bool fold) {
    const uint32_t KS = (S + 3) / 4, K = KS * 32, SP = KS * 4;
    const size_t b_bytes = (size_t)K * tc_total_cols(T, fold);
    return ((b_bytes + 127) & ~(size_t)127) + (size_t)kTcGroups * 128 * K + (size_t)(8 * SP + 11 * T) * sizeof(u64) + kTcGroups * 8 + 16;
}

int lincomb_tc_launch(fhe_b200_lincomb* lc, const LcView& view, uint32_t n, uint32_t batch, cudaStream_t st) {
    LcTcArgs a;
    a.src_mod = lc->src_mod; a.pre = lc->pre; a.pre_s = lc->pre_s; a.th_hi = lc->th_hi; a.th_lo = lc->th_lo;
    a.dst_mod = lc->dst_mod; a.mu_hi = lc->mu_hi; a.mu_lo = lc->mu_lo; a.c = lc->c; a.lam = lc->lam;
    a.bmat = lc->d_tc_b; a.fold = lc->d_tc_fold;
    a.v = view;
    a.S = lc->S; a.T = lc->T; a.KS = (lc->S + 3) / 4; a.logn = host::ilog2(n);
    a.use_pre = lc->use_pre; a.use_extra = lc->use_extra; a.c_is_one = lc->use_pre ? 0 : 1;
    a.tiles = (size_t)batch * n / 128;
    FHE_REQUIRE(!(lc->tc_fold && view.sub && !view.epi_scalar_shoup), "lincomb (tcgen05 path): the fused epilogue needs epi_scalar_shoup");
    const size_t smem = lincomb_tc_smem_bytes(lc->S, lc->T, lc->tc_fold);
    static size_t attr_smem[64] = {0};
    if (smem > attr_smem[lc->device & 63]) {
#define TC_ATTR(M_, K_) FHE_CUDA(cudaFuncSetAttribute(lincomb_tc_kernel<M_, K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TC_ATTR(false, 0) TC_ATTR(false, 1) TC_ATTR(false, 2) TC_ATTR(false, 3) TC_ATTR(true, 0) TC_ATTR(true, 1) TC_ATTR(true, 2) TC_ATTR(true, 3)
#undef TC_ATTR
        attr_smem[lc->device & 63] = smem;
    }
    const unsigned grid = (unsigned)(a.tiles < (size_t)lc->sm_count ? a.tiles : (size_t)lc->sm_count);
    const int kind = (!a.use_extra && !a.c_is_one && !view.sub) ? 0 : (a.use_extra && a.c_is_one && !view.sub) ? 1
                   : (!a.use_extra && !a.c_is_one && view.sub) ? 2 : 3;
#define TC_GO(M_, K_) lincomb_tc_kernel<M_, K_><<<grid, 128 * kTcGroups, smem, st>>>(a)
    if (lc->tc_fold) { if (kind == 0) TC_GO(true, 0); else if (kind == 1) TC_GO(true, 1); else if (kind == 2) TC_GO(true, 2); else TC_GO(true, 3); }
    else { if (kind == 0) TC_GO(false, 0); else if (kind == 1) TC_GO(false, 1); else if (kind == 2) TC_GO(false, 2); else TC_GO(false, 3); }
#undef TC_GO
    FHE_LAUNCH_CHECK();
    return 0;
}

}  // namespace fhe_b200
