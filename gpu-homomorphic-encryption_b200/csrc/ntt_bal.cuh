// Balanced two-pass negacyclic NTT bodies for 2^13 <= N <= 2^16 (host+device so tests/emul can run them on the CPU).
//
// Same role as ntt_core.cuh (replaces /root/reference/kernels/ntt_kernels.cu:7-121, one butterfly per thread per stage on a
// single block), different split of the log N stages: the integer (IMAD) pipe binds this transform, so the two passes are
// given equal arithmetic and the least possible work besides butterflies -- two radix-16 register rounds per pass and one
// exchange through shared memory between them.
//
//   pass A  "column pass": the first KA = logN - 8 stages (strides N/2 .. 256).  The limb is a 2^KA x 256 matrix; an item is
//            all rows x C adjacent columns (4096 elements, C = 4096 >> KA) and a CTA of 256 threads runs it as
//            round 1 (R1 = KA - 4 stages on the high row bits; a thread holds 2^R1 rows x 2^(4-R1) columns), exchange,
//            round 2 (4 stages on the low row bits).  Twiddles: the limb's first 2^KA table entries, staged once per CTA.
//   pass B  "tile pass":   the last 8 stages on contiguous 256-element tiles.  A WARP owns a pair of adjacent tiles (4 KiB):
//            round 1 (strides 128..16; lane = (tile, low nibble)), exchange inside the warp's own 4 KiB of shared memory,
//            round 2 (strides 8..1; lane = (tile, high nibble), 16 contiguous elements), coalesced 16-byte copy-out.
//            No block barrier anywhere in this pass.  Twiddles: one 8 KiB block per tile pair (layout below), staged by one
//            bulk copy and reused for every polynomial the warp processes.
// The inverse runs the mirror image (B' then A').  Butterflies, lazy bounds and twiddle indexing are those of ntt_core.cuh.
#pragma once
#include "ntt_core.cuh"

namespace fhe_b200 {

FHE_HDC int bal_ka(int logn) { return logn - 8; }
FHE_HDC bool bal_supported(int logn) { return logn >= 13 && logn <= 16; }
// entries of one tile pair's staged twiddle block (pass B)
FHE_HDC int bal_pair_entries() { return 512; }

// ---- twiddle accessors ------------------------------------------------------------------------------------------------
// pass A round 1: stage v, group = low v bits of key (the bits of key above v select the thread's column, not a group)
struct TwA1 {
    const Twiddle* s;
    FHE_HD Twiddle get(int v, int key) const { return ld_tw_s(s + ((1 << v) + (key & ((1 << v) - 1)))); }
};
// pass A round 2: root = 2^R1 + rh
struct TwA2 {
    const Twiddle* s; u32 root;
    FHE_HD Twiddle get(int v, int key) const { return ld_tw_s(s + ((root << v) + key)); }
};
// Staged block of one tile pair (tiles 2p, 2p+1 of a limb; root_b = 2^KA + b):
//   [t*16 + (1<<v)-1 + key]              round 1 of tile t (stage KA+v, key < 2^v): main[(root_b << v) + key]      (2 x 15, entries 15 and 31 unused)
//   [32 + 32*((1<<v)-1) + key*32 + lane] round 2 (stage KA+4+v), lane = (t, E):     main[((root_b*16 + E) << v) + key]  (15 x 32)
struct TwB1 {
    const Twiddle* s;            // already offset by t*16
    FHE_HD Twiddle get(int v, int key) const { return ld_tw_s(s + ((1 << v) - 1 + key)); }
};
struct TwB2 {
    const Twiddle* s; u32 lane;
    FHE_HD Twiddle get(int v, int key) const { return ld_tw_s(s + (32 + 32 * ((1 << v) - 1) + key * 32 + lane)); }
};

// =======================================================================================================================
// pass A.  g = limb-polynomial base + first column of the item (row stride 256 elements); s = 4096 u64 of shared memory;
// stw = the limb's first 2^KA twiddles (forward or inverse table).
// =======================================================================================================================
template <int KA, int HB, bool NEAR>
struct BalA {
    static_assert(KA >= 5 && KA <= 8, "pass A covers 5..8 stages");
    static constexpr int R1 = KA - 4;
    static constexpr u32 RM = (1u << R1) - 1;
    static constexpr int C = 4096 >> KA;             // columns per item
    static constexpr int CB = 256 / C;               // items (column blocks) per limb-polynomial
    static FHE_HDC int fwd_out_bound() { return fwd_bound_after(1, KA, HB, NEAR); }

    // round-1 thread (rl, ct) holds element e = (cc << R1) | rh: row rh*16 + rl, column cc*16 + ct
    struct Off1 { static FHE_HDC long at(int e) { return (long)((e & (int)RM) << 4) * 256 + ((e >> R1) << 4); } };       // at1(tid, e) - at1(tid, 0)
    static FHE_HD size_t at1(u32 tid, int e) { return (size_t)(((e & RM) << 4) | (tid >> 4)) * 256 + (((u32)e >> R1) << 4) + (tid & 15); }
    // round-2 thread (e1, ct), e1 = (cc << R1) | rh, holds rows rh*16 + rl
    static FHE_HD size_t at2(u32 tid, int rl) { const u32 e1 = tid >> 4; return (size_t)(((e1 & RM) << 4) | (u32)rl) * 256 + ((e1 >> R1) << 4) + (tid & 15); }

    static FHE_HD void fwd_round1(u32 tid, const u64* g, u64* s, const Twiddle* stw, const LimbParams& P) {
        u64 x[16];
        ldg16o<Off1>(g + (size_t)(tid >> 4) * 256 + (tid & 15), x);
        fwd_stages<4, R1, HB, NEAR, 1>(x, TwA1{stw}, P);
#pragma unroll
        for (int e = 0; e < 16; e++) s[(e << 8) | tid] = x[e];
    }
    static FHE_HD void fwd_round2(u32 tid, u64* g, const u64* s, const Twiddle* stw, const LimbParams& P) {
        u64 x[16];
        const u32 e1 = tid >> 4, ct = tid & 15;
#pragma unroll
        for (int rl = 0; rl < 16; rl++) x[rl] = s[(e1 << 8) | (rl << 4) | ct];
        fwd_stages<4, 4, HB, NEAR, fwd_bound_after(1, R1, HB, NEAR)>(x, TwA2{stw, (1u << R1) + (e1 & RM)}, P);
        stg16o<Stride16::Of<256>>(g + at2(tid, 0), x);
    }

    // inverse: round 2' (low row bits) first, entry bound BIN = what pass B' leaves; then round 1' ends the transform (N^-1 folded)
    template <int BIN>
    static FHE_HD void inv_round2(u32 tid, const u64* g, u64* s, const Twiddle* stw, const LimbParams& P) {
        u64 x[16];
        const u32 e1 = tid >> 4, ct = tid & 15;
        ldg16<256>(g + at2(tid, 0), x);
        inv_round16<4, HB, NEAR, false, BIN, kInvCap3>(x, TwA2{stw, (1u << R1) + (e1 & RM)}, P);
#pragma unroll
        for (int rl = 0; rl < 16; rl++) s[(e1 << 8) | (rl << 4) | ct] = x[rl];
    }
    template <int BIN>
    static FHE_HD void inv_round1(u32 tid, u64* g, const u64* s, const Twiddle* stw, const LimbParams& P) {
        u64 x[16];
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = s[(e << 8) | tid];
        inv_round16<R1, HB, NEAR, true, inv_round16_out(4, HB, NEAR, false, BIN, kInvCap3), HB / 2>(x, TwA1{stw}, P);
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = normalize_last<HB, NEAR, R1>(x[e], e, P.q);
        stg16o<Off1>(g + at1(tid, 0), x);
    }
    // as inv_round1, but every output goes through put(row, column inside the item, value): row of the limb's 2^KA x 256 matrix.
    // Used by the scatter form of the last inverse pass (limb-sharded execution): the rows of a coefficient block belong to the
    // GPU that owns that block, and the store goes straight to its buffer.
    template <int BIN, class PUT>
    static FHE_HD void inv_round1_put(u32 tid, const PUT& put, const u64* s, const Twiddle* stw, const LimbParams& P) {
        u64 x[16];
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = s[(e << 8) | tid];
        inv_round16<R1, HB, NEAR, true, inv_round16_out(4, HB, NEAR, false, BIN, kInvCap3), HB / 2>(x, TwA1{stw}, P);
#pragma unroll
        for (int e = 0; e < 16; e++)
            put((((u32)e & RM) << 4) | (tid >> 4), (((u32)e >> R1) << 4) + (tid & 15), normalize_last<HB, NEAR, R1>(x[e], e, P.q));
    }
};

// =======================================================================================================================
// pass B.  g = base of the tile pair (512 contiguous elements); s = the warp's 512 u64 of shared memory (swizzled rows of
// 16 elements, see swz()); sb = the pair's staged twiddle block.  Phases are separated by __syncwarp() in the kernel.
// =======================================================================================================================
template <int HB, bool NEAR>
struct BalB {
    // logical 16-byte chunk c (0..255) of the pair <-> its swizzled position
    static FHE_HD u32 chunk_pos(u32 c) { return ((c >> 3) << 3) | ((c & 7) ^ ((c >> 3) & 7)); }

    // ---- forward: B0 = bound left by pass A
    // phase 1 is split so that the kernel can issue the global loads early (before the staged twiddles have landed, and for
    // the next polynomial while the current one is copied out)
    static FHE_HD void fwd_load(u32 lane, const u64* g, u64 (&x)[16]) {
        const u32 t = lane >> 4, j = lane & 15;
        ldg16<16>(g + ((t << 8) | j), x);
    }
    // the same sixteen values from a tile pair that a bulk copy has landed in shared memory in natural order (st)
    static FHE_HD void fwd_load_staged(u32 lane, const u64* st, u64 (&x)[16]) {
        const u32 t = lane >> 4, j = lane & 15;
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = st[(t << 8) | (e << 4) | j];
    }
    template <int B0>
    static FHE_HD void fwd_phase1(u32 lane, u64 (&x)[16], u64* s, const Twiddle* sb, const LimbParams& P) {
        const u32 t = lane >> 4, j = lane & 15;
        fwd_stages<4, 4, HB, NEAR, B0>(x, TwB1{sb + t * 16}, P);
#pragma unroll
        for (int e = 0; e < 16; e++) s[swz((t << 8) | (e << 4) | j)] = x[e];
    }
    // forward round 1 on a tile pair that a bulk copy has landed in s in natural order; the exchange layout overwrites it in place
    // (every lane has read its sixteen inputs before any lane stores: the warp barrier in the middle)
    template <int B0>
    static FHE_HD void fwd_phase1_inplace(u32 lane, u64* s, const Twiddle* sb, const LimbParams& P) {
        const u32 t = lane >> 4, j = lane & 15;
        u64 x[16];
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = s[(t << 8) | (e << 4) | j];
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
        fwd_stages<4, 4, HB, NEAR, B0>(x, TwB1{sb + t * 16}, P);
#pragma unroll
        for (int e = 0; e < 16; e++) s[swz((t << 8) | (e << 4) | j)] = x[e];
    }
    // LAZY: the results stay in [0, fwd_lazy_bound() q) instead of being brought into [0, q) -- for consumers that multiply them with a
    // 128-bit Barrett reduction anyway (tensor product, key-switch inner product): a range reduction and a conditional subtraction
    // per element less
    template <int B0, bool LAZY = false>
    static FHE_HD void fwd_phase2(u32 lane, u64* s, const Twiddle* sb, const LimbParams& P) {
        u64 x[16];
        const u64 q = P.q;
        const u32 row = lane << 4;
#pragma unroll
        for (int k = 0; k < 8; k++) ld2(s + (row | ((k ^ (lane & 7)) << 1)), x[2 * k], x[2 * k + 1]);
        constexpr int B1 = fwd_bound_after(B0, 4, HB, NEAR);
        constexpr int BE = fwd_bound_after(B1, 4, HB, NEAR);
        fwd_stages<4, 4, HB, NEAR, B1>(x, TwB2{sb, lane}, P);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (LAZY) st2s(s + (row | ((k ^ (lane & 7)) << 1)), x[2 * k], x[2 * k + 1]);
            else st2s(s + (row | ((k ^ (lane & 7)) << 1)), normalize<HB, NEAR, BE>(x[2 * k], q), normalize<HB, NEAR, BE>(x[2 * k + 1], q));
        }
    }
    template <int B0>
    static FHE_HDC int fwd_lazy_bound() { return fwd_bound_after(fwd_bound_after(B0, 4, HB, NEAR), 4, HB, NEAR); }
    // forward round 2 with the sixteen canonical results left in registers (the lane's own row of the exchange buffer in,
    // nothing written): the fused tile kernels (ntt_fused.cu) go on to the pointwise work from here
    template <int B0>
    static FHE_HD void fwd_phase2_regs(u32 lane, const u64* s, const Twiddle* sb, const LimbParams& P, u64 (&x)[16]) {
        const u64 q = P.q;
        row_load(lane, s, x);
        constexpr int B1 = fwd_bound_after(B0, 4, HB, NEAR);
        constexpr int BE = fwd_bound_after(B1, 4, HB, NEAR);
        fwd_stages<4, 4, HB, NEAR, B1>(x, TwB2{sb, lane}, P);
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = normalize<HB, NEAR, BE>(x[e], q);
    }
    // the lane's own row (16 consecutive elements of the tile pair: lane * 16 ..) of a warp buffer
    static FHE_HD void row_load(u32 lane, const u64* s, u64 (&x)[16]) {
        const u32 row = lane << 4;
#pragma unroll
        for (int k = 0; k < 8; k++) ld2(s + (row | ((k ^ (lane & 7)) << 1)), x[2 * k], x[2 * k + 1]);
    }
    static FHE_HD void row_store(u32 lane, u64* s, const u64 (&x)[16]) {
        const u32 row = lane << 4;
#pragma unroll
        for (int k = 0; k < 8; k++) st2(s + (row | ((k ^ (lane & 7)) << 1)), x[2 * k], x[2 * k + 1]);
    }
    // chunk k (elements 2k, 2k+1) of the lane's own row
    static FHE_HD u64* row_chunk(u32 lane, u64* s, int k) { return s + ((lane << 4) | (((u32)k ^ (lane & 7)) << 1)); }
    static FHE_HD void fwd_phase3(u32 lane, u64* g, const u64* s) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 c = lane + 32 * i;
            u64 a, b; ld2(s + 2 * chunk_pos(c), a, b); st2(g + 2 * c, a, b);
        }
    }

    // ---- inverse
    static FHE_HD void inv_phase1(u32 lane, const u64* g, u64* s) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 c = lane + 32 * i;
            u64 a, b; ldg2(g + 2 * c, a, b); st2(s + 2 * chunk_pos(c), a, b);
        }
    }
    // the same re-layout from a tile pair that a bulk copy has landed in shared memory in natural order (st)
    static FHE_HD void inv_phase1_staged(u32 lane, const u64* st, u64* s) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 c = lane + 32 * i;
            u64 a, b; ld2(st + 2 * c, a, b); st2(s + 2 * chunk_pos(c), a, b);
        }
    }
    // inverse round 1 on a row that is already in registers (canonical values); the result goes to the lane's own row of s
    static FHE_HD void inv_phase2_regs(u32 lane, u64 (&x)[16], u64* s, const Twiddle* sb, const LimbParams& P) {
        inv_round16<4, HB, NEAR, false, 1, kInvCap1>(x, TwB2{sb, lane}, P);
        row_store(lane, s, x);
    }
    static FHE_HD void inv_phase2(u32 lane, u64* s, const Twiddle* sb, const LimbParams& P) {
        u64 x[16];
        const u32 row = lane << 4;
#pragma unroll
        for (int k = 0; k < 8; k++) ld2(s + (row | ((k ^ (lane & 7)) << 1)), x[2 * k], x[2 * k + 1]);
        inv_round16<4, HB, NEAR, false, 1, kInvCap1>(x, TwB2{sb, lane}, P);
#pragma unroll
        for (int k = 0; k < 8; k++) st2s(s + (row | ((k ^ (lane & 7)) << 1)), x[2 * k], x[2 * k + 1]);
    }
    static FHE_HD void inv_phase3(u32 lane, u64* g, const u64* s, const Twiddle* sb, const LimbParams& P) {
        const u32 t = lane >> 4, j = lane & 15;
        u64 x[16];
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = s[swz((t << 8) | (e << 4) | j)];
        inv_round16<4, HB, NEAR, false, inv_round16_out(4, HB, NEAR, false, 1, kInvCap1), kInvCap2>(x, TwB1{sb + t * 16}, P);
#pragma unroll
        for (int e = 0; e < 16; e++) g[(t << 8) | (e << 4) | j] = x[e];
    }
    static FHE_HDC int inv_out_bound() { return inv_round16_out(4, HB, NEAR, false, inv_round16_out(4, HB, NEAR, false, 1, kInvCap1), kInvCap2); }
};

}  // namespace fhe_b200
