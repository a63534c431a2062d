// Negacyclic NTT bodies for the B200 engine (host+device so tests/emul can run them on the CPU).
//
// Replaces /root/reference/kernels/ntt_kernels.cu:7-121 (ntt_forward/inverse_optimized_kernel, one
// butterfly per thread per stage, 256-bit Montgomery, single block) and :140-161 (bit_reverse_kernel,
// eliminated: forward leaves X[k] at position bitrev(k), inverse consumes that order).
//
// Decomposition of an N = 2^logN point transform of one limb (one modulus):
//   pass A  "row pass"   : the first K1 = logN - LB stages (strides N/2 .. NB).  Registers only:
//                          a thread owns V adjacent columns x 2^K1 rows (row stride NB = 2^LB elements).
//   pass B  "tile pass"  : the remaining LB stages on each contiguous tile of NB elements, in shared
//                          memory, as three register rounds of radix 16 / 16 / 2^(LB-8).
//   N <= 4096 uses pass B alone (K1 = 0).  The inverse runs the mirror image (B' then A').
//
// Twiddle index of the butterfly group g at stage s (m = 2^s groups) is m + g in the bit-reversed
// psi-power table (Twiddle{w, floor(w 2^64 / q)}); a tile with index b starts from "root" 2^K1 + b and every
// sub-transform a thread performs is addressed as (root << v) + key.
//
// Lazy arithmetic: values live in [0, B*q) with the bound B tracked at compile time (template recursion over the
// stages).  HB = largest power of two with HB*q_max <= 2^64 is the head-room (16 for q < 2^60, 8 for q < 2^61).
// The twiddle product T = shoup_mul_lazy3(.) lies in [0, kTQ q) = [0, 3q) for any 64-bit input.
//   forward (CT):  X' = X + T, Y' = X + 3q - T: the bound grows by 3 per stage; X gets one range reduction only in the
//                  stages where B + 3 would exceed HB.
//   inverse (GS):  S = X + Y, Y' = T(X + Bq - Y): the sum bound doubles; S gets one conditional subtraction of B q once
//                  doubling again would overflow (steady state B = HB/2).
// The conditional subtractions run on the ALU pipe, which has slack; the FMA pipe (IMAD) is the binding one.
#pragma once
#include "modarith.cuh"

// host-only bound checking for the CPU emulation tests (compile the emulator with -DFHE_CHECK_BOUNDS)
#if defined(FHE_CHECK_BOUNDS) && !defined(__CUDA_ARCH__)
#include <cstdio>
#include <cstdlib>
#define FHE_BOUND(val, B, q)                                                                                      \
    do {                                                                                                          \
        if ((unsigned __int128)(val) >= (unsigned __int128)(B) * (q)) {                                           \
            std::fprintf(stderr, "lazy bound violated: %s >= %d*q at %s:%d\n", #val, (int)(B), __FILE__, __LINE__); \
            std::abort();                                                                                         \
        }                                                                                                         \
    } while (0)
#else
#define FHE_BOUND(val, B, q) do { } while (0)
#endif

namespace fhe_b200 {

FHE_HD Twiddle ld_tw(const Twiddle* p) {
#if defined(__CUDA_ARCH__)
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
    Twiddle t; t.w = v.x; t.ws = v.y; return t;
#else
    return *p;
#endif
}

// shared-memory swizzle: 128-byte rows of 16 elements, 16-byte chunk index XOR (row & 7)
FHE_HD u32 swz(u32 idx) { return idx ^ (((idx >> 4) & 7u) << 1); }

// 16-byte (two element) accesses; p must be 16-byte aligned
FHE_HD void ld2(const u64* p, u64& a, u64& b) {
#if defined(__CUDA_ARCH__)
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
    a = v.x; b = v.y;
#else
    a = p[0]; b = p[1];
#endif
}
FHE_HD void st2(u64* p, u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a, b);
#else
    p[0] = a; p[1] = b;
#endif
}
// the same two elements as two 8-byte stores: a 16-byte store wants its four words in an aligned register quad, and for values that
// come out of three-input adds or selects ptxas gathers them there with moves -- which it issues as IMAD.MOV on the binding pipe
FHE_HD void st2s(u64* p, u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    asm volatile("st.volatile.shared.u64 [%0], %1;\n\tst.volatile.shared.u64 [%0+8], %2;" :: "r"((u32)__cvta_generic_to_shared(p)), "l"(a), "l"(b) : "memory");
#else
    st2(p, a, b);
#endif
}

// Global loads of polynomial data go to L2 only (ld.global.cg): the data is streamed (no L1 reuse), and in the fused
// row+tile kernel another SM may have rewritten a line that this SM's L1 still holds from the row pass.
FHE_HD u64 ldg1(const u64* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
FHE_HD void ldg2(const u64* p, u64& a, u64& b) {
#if defined(__CUDA_ARCH__)
    const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(p));
    a = v.x; b = v.y;
#else
    a = p[0]; b = p[1];
#endif
}

// Sixteen L2-only loads base[OFF::at(e)], e = 0..15 (compile-time element offsets), as ONE asm statement: the compiler may not sink some of them below the
// first butterflies (it did, to save registers: four of the sixteen loads of the tile pass were issued 200 instructions late and
// `long_scoreboard` became the top stall, 0.78 ms against 0.65 ms per pass at config 3).
struct Stride16 { template <int STRIDE> struct Of { static FHE_HDC long at(int e) { return (long)e * STRIDE; } }; };
template <class OFF>
FHE_HD void ldg16o(const u64* base, u64 (&x)[16]) {
#if defined(__CUDA_ARCH__)
    asm volatile(
        "ld.global.cg.u64 %0, [%16+%17];\n\t"  "ld.global.cg.u64 %1, [%16+%18];\n\t"  "ld.global.cg.u64 %2, [%16+%19];\n\t"
        "ld.global.cg.u64 %3, [%16+%20];\n\t"  "ld.global.cg.u64 %4, [%16+%21];\n\t"  "ld.global.cg.u64 %5, [%16+%22];\n\t"
        "ld.global.cg.u64 %6, [%16+%23];\n\t"  "ld.global.cg.u64 %7, [%16+%24];\n\t"  "ld.global.cg.u64 %8, [%16+%25];\n\t"
        "ld.global.cg.u64 %9, [%16+%26];\n\t"  "ld.global.cg.u64 %10, [%16+%27];\n\t" "ld.global.cg.u64 %11, [%16+%28];\n\t"
        "ld.global.cg.u64 %12, [%16+%29];\n\t" "ld.global.cg.u64 %13, [%16+%30];\n\t" "ld.global.cg.u64 %14, [%16+%31];\n\t"
        "ld.global.cg.u64 %15, [%16+%32];"
        : "=l"(x[0]), "=l"(x[1]), "=l"(x[2]), "=l"(x[3]), "=l"(x[4]), "=l"(x[5]), "=l"(x[6]), "=l"(x[7]), "=l"(x[8]), "=l"(x[9]),
          "=l"(x[10]), "=l"(x[11]), "=l"(x[12]), "=l"(x[13]), "=l"(x[14]), "=l"(x[15])
        : "l"(base), "n"(OFF::at(0) * 8), "n"(OFF::at(1) * 8), "n"(OFF::at(2) * 8), "n"(OFF::at(3) * 8), "n"(OFF::at(4) * 8),
          "n"(OFF::at(5) * 8), "n"(OFF::at(6) * 8), "n"(OFF::at(7) * 8), "n"(OFF::at(8) * 8), "n"(OFF::at(9) * 8), "n"(OFF::at(10) * 8),
          "n"(OFF::at(11) * 8), "n"(OFF::at(12) * 8), "n"(OFF::at(13) * 8), "n"(OFF::at(14) * 8), "n"(OFF::at(15) * 8)
        : "memory");
#else
    for (int e = 0; e < 16; e++) x[e] = base[OFF::at(e)];
#endif
}
template <int STRIDE>
FHE_HD void ldg16(const u64* base, u64 (&x)[16]) { ldg16o<Stride16::Of<STRIDE>>(base, x); }

// Sixteen stores base[OFF::at(e)] = x[e] with the element offsets as immediates of ONE base register pair: written as g[f(tid, e)]
// the compiler rebuilds a 64-bit address per store (LOP3, IADD3, IADD3.X, LEA, LEA.HI.X: two issue cycles per butterfly of the column pass)
template <class OFF>
FHE_HD void stg16o(u64* base, const u64 (&x)[16]) {
#if defined(__CUDA_ARCH__)
    asm volatile(
        "st.global.u64 [%16+%17], %0;\n\t"  "st.global.u64 [%16+%18], %1;\n\t"  "st.global.u64 [%16+%19], %2;\n\t"
        "st.global.u64 [%16+%20], %3;\n\t"  "st.global.u64 [%16+%21], %4;\n\t"  "st.global.u64 [%16+%22], %5;\n\t"
        "st.global.u64 [%16+%23], %6;\n\t"  "st.global.u64 [%16+%24], %7;\n\t"  "st.global.u64 [%16+%25], %8;\n\t"
        "st.global.u64 [%16+%26], %9;\n\t"  "st.global.u64 [%16+%27], %10;\n\t" "st.global.u64 [%16+%28], %11;\n\t"
        "st.global.u64 [%16+%29], %12;\n\t" "st.global.u64 [%16+%30], %13;\n\t" "st.global.u64 [%16+%31], %14;\n\t"
        "st.global.u64 [%16+%32], %15;"
        :: "l"(x[0]), "l"(x[1]), "l"(x[2]), "l"(x[3]), "l"(x[4]), "l"(x[5]), "l"(x[6]), "l"(x[7]), "l"(x[8]), "l"(x[9]),
           "l"(x[10]), "l"(x[11]), "l"(x[12]), "l"(x[13]), "l"(x[14]), "l"(x[15]),
           "l"(base), "n"(OFF::at(0) * 8), "n"(OFF::at(1) * 8), "n"(OFF::at(2) * 8), "n"(OFF::at(3) * 8), "n"(OFF::at(4) * 8),
           "n"(OFF::at(5) * 8), "n"(OFF::at(6) * 8), "n"(OFF::at(7) * 8), "n"(OFF::at(8) * 8), "n"(OFF::at(9) * 8), "n"(OFF::at(10) * 8),
           "n"(OFF::at(11) * 8), "n"(OFF::at(12) * 8), "n"(OFF::at(13) * 8), "n"(OFF::at(14) * 8), "n"(OFF::at(15) * 8)
        : "memory");
#else
    for (int e = 0; e < 16; e++) base[OFF::at(e)] = x[e];
#endif
}

// bounds in units of q; the twiddle product is lazy in [0, kTQ q) (shoup_mul_lazy3).
// NEAR: every modulus q satisfies 2^60 - 2^32 < q < 2^60, so near60_reduce brings anything below 16q under 2q.
FHE_HDC int red_to(int HB, bool NEAR) { return NEAR ? 2 : HB / 2; }
FHE_HDC int fwd_bound_after(int B, int stages, int HB, bool NEAR) {
    for (int i = 0; i < stages; i++) { if (B + kTQ > HB) B = red_to(HB, NEAR); B += kTQ; }
    return B;
}
FHE_HDC int inv_bound_step(int B, int HB, bool NEAR) {
    int nb = 2 * B > kTQ ? 2 * B : kTQ;
    if (2 * nb > HB) { const int r = NEAR ? 2 : nb / 2; nb = r > kTQ ? r : kTQ; }
    return nb;
}
FHE_HDC int inv_bound_after(int B, int stages, int HB, bool NEAR) {
    for (int i = 0; i < stages; i++) B = inv_bound_step(B, HB, NEAR);
    return B;
}

// ---- twiddle accessors: where stage v / group key of the current sub-transform finds its twiddle ------------------
// global table, SEAL-style index (root << (SH0+v)) + key
template <int SH0>
struct TwGlobal {
    const Twiddle* tw; u32 root;
    FHE_HD Twiddle get(int v, int key) const { return ld_tw(tw + ((root << (SH0 + v)) + key)); }
};
FHE_HD Twiddle ld_tw_s(const Twiddle* p) {
#if defined(__CUDA_ARCH__)
    const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(p);
    Twiddle w; w.w = t.x; w.ws = t.y; return w;
#else
    return *p;
#endif
}
// Per-tile staged tables (shared memory on the device).  Layout of one tile's block "p12" (256 entries):
//   phase 1 (first 4 tile stages, uniform over the CTA):   [ (1<<v)-1 + key ]                       v<4, key<2^v   (15)
//   phase 2 (next 4, depends on hi = tid >> (LB-8) < 16):  [ 15 + 16*((1<<v)-1) + (hi<<v) + key ]                 (240)
// and of the block "p3" (last R3 = LB-8 stages, every thread its own twiddles), stored [stage][key][tid] so that
// consecutive lanes read consecutive 16-byte entries:        [ NT*KB*((1<<v)-1) + key*NT + tid ],  KB = 2^(4-R3), key < KB*2^v
struct TwP1 {
    const Twiddle* s;
    FHE_HD Twiddle get(int v, int key) const { return ld_tw_s(s + ((1 << v) - 1 + key)); }
};
struct TwP2 {
    const Twiddle* s; u32 hi;
    FHE_HD Twiddle get(int v, int key) const { return ld_tw_s(s + (15 + 16 * ((1 << v) - 1) + (hi << v) + key)); }
};
template <int NT, int R3>
struct TwP3 {
    const Twiddle* s; u32 tid;
    FHE_HD Twiddle get(int v, int key) const { return ld_tw_s(s + (NT * (1 << (4 - R3)) * ((1 << v) - 1) + key * NT + tid)); }
};
// sizes (entries) of the staged blocks
FHE_HDC int p12_entries() { return 256; }
FHE_HDC int p3_entries(int LB) { return (1 << (LB - 4)) * (16 - (1 << (12 - LB))); }

// ---- forward stages on a register array of E = 2^LE elements: R stages, active bits = low R bits of e ----
// B (template) is the bound on entry in units of q; the bound on exit is fwd_bound_after(B, R, HB).
// Compile-time recursion over the stage index V keeps every array index and every bound a constant.
template <int LE, int R, int HB, bool NEAR, int B, class TW, int V = 0>
FHE_HD void fwd_stages(u64 (&x)[1 << LE], const TW& tw, const LimbParams& P) {
    if constexpr (V < R) {
        constexpr int st = 1 << (R - 1 - V);
        constexpr int sh = LE - R + V;
        constexpr bool red = (B + kTQ > HB);
        const u64 q = P.q, tq = P.tq, nq = 0 - q;
        const u64 hq = (u64)(HB / 2) * q;
#pragma unroll
        for (int key = 0; key < (1 << sh); key++) {
            const Twiddle w = tw.get(V, key);
#pragma unroll
            for (int j = 0; j < st; j++) {
                const int e = (key << (R - V)) | j;
                u64 X = x[e];
                FHE_BOUND(X, B, q); FHE_BOUND(x[e + st], B, q);
                if (red) X = NEAR ? near60_reduce(X, nq) : csub(X, hq);
                const u64 T = shoup_mul_lazy3(x[e + st], w.w, w.ws, nq);
                FHE_BOUND(T, kTQ, q); FHE_BOUND((unsigned __int128)X + T, (red ? red_to(HB, NEAR) : B) + kTQ, q);
                x[e] = add3z(X, T);
                x[e + st] = X + tq - T;
            }
        }
        fwd_stages<LE, R, HB, NEAR, (red ? red_to(HB, NEAR) : B) + kTQ, TW, V + 1>(x, tw, P);
    }
}

// ---- inverse stages (mirror order: V runs R-1 .. 0).  LAST: the final stage of the whole transform folds N^-1 in.
// exit bound: inv_bound_after(B, R, HB), or kTQ when LAST.
template <int LE, int R, int HB, bool NEAR, bool LAST, int B, class TW, int V = R - 1>
FHE_HD void inv_stages(u64 (&x)[1 << LE], const TW& tw, const LimbParams& P) {
    if constexpr (V >= 0) {
        constexpr int st = 1 << (R - 1 - V);
        constexpr int sh = LE - R + V;
        constexpr bool last = LAST && (V == 0);
        constexpr int SB = 2 * B > kTQ ? 2 * B : kTQ;          // bound after this stage before any reduction
        constexpr bool red = !last && (2 * SB > HB);           // reduce the sums so that the next stage cannot overflow
        static_assert(2 * B <= HB, "inverse stage would overflow 64 bits");
        const u64 q = P.q, nq = 0 - q;
        const u64 bq = (B % kTQ == 0) ? (u64)(B / kTQ) * P.tq : (u64)B * q;      // multiples of 3q from the loaded 3q (no multiply by 3 in the loop)
#pragma unroll
        for (int key = 0; key < (1 << sh); key++) {
            Twiddle w;
            if (last) { w.w = P.w1ninv; w.ws = P.w1ninv_s; }
            else w = tw.get(V, key);
#pragma unroll
            for (int j = 0; j < st; j++) {
                const int e = (key << (R - V)) | j;
                const u64 X = x[e], Y = x[e + st];
                FHE_BOUND(X, B, q); FHE_BOUND(Y, B, q); FHE_BOUND((unsigned __int128)X + Y, HB, q);
                u64 S = add3z(X, Y);
                const u64 D = X + bq - Y;
                if (last) S = scale_ninv(S, P.nm, q);
                else if (red) S = NEAR ? near60_reduce(S, nq) : csub(S, bq);
                x[e] = S;
                x[e + st] = shoup_mul_lazy3(D, w.w, w.ws, nq);
            }
        }
        inv_stages<LE, R, HB, NEAR, LAST, inv_bound_step(B, HB, NEAR), TW, V - 1>(x, tw, P);
    }
}

// ---- inverse stages on sixteen values with PER-ELEMENT lazy bounds (moduli just below 2^60) ---------------------------------------
// inv_stages tracks one bound for all values, so every sum is range-reduced every other stage (T < 3q: 3 -> 6 -> 12 -> reduce).  But a
// value's bound is decided by its own history -- 3 after a twiddle product, the sum of its two inputs' bounds after an addition, and the
// two inputs of a butterfly always share their history -- and the history is the index: after the stage with stride st, element e holds a
// sum if bit st of e is clear, a product if it is set.  Tracking the sixteen bounds separately (four bits each, packed in a 64-bit
// template constant) only the sums that would pass HB/2 are reduced: 40 instead of 56 reductions per sixteen values per 2^16-point
// transform, 18 instead of 32 in the inverse column pass.  Bounds become uniform again where the values change threads: the last stage
// of a round reduces what is above CAP, and the next round starts from the largest bound left.
FHE_HDC int bs_get(u64 bs, int e) { return (int)((bs >> (4 * e)) & 15u) + 1; }
FHE_HDC u64 bs_put(u64 bs, int e, int b) { return (bs & ~(15ull << (4 * e))) | ((u64)(b - 1) << (4 * e)); }
FHE_HDC u64 bs_all(int b) { u64 r = 0; for (int e = 0; e < 16; e++) r = bs_put(r, e, b); return r; }
FHE_HDC int bs_max(u64 bs) { int m = 1; for (int e = 0; e < 16; e++) m = bs_get(bs, e) > m ? bs_get(bs, e) : m; return m; }
// bounds after one Gentleman-Sande stage with stride st: sums above `limit` are range-reduced (near60_reduce: below 2q); `scaled`: the
// last stage of the whole transform, whose sums go through scale_ninv (below 2q)
FHE_HDC u64 bs_gs_stage(u64 bs, int st, int limit, bool scaled) {
    u64 r = bs;
    for (int e = 0; e < 16; e++)
        if (!(e & st)) {
            int sum = bs_get(bs, e) + bs_get(bs, e + st);
            if (scaled || sum > limit) sum = 2;
            r = bs_put(r, e, sum); r = bs_put(r, e + st, kTQ);
        }
    return r;
}
FHE_HDC u64 bs_gs_round(u64 bs, int R, int HB, int cap, bool last) {
    for (int V = R - 1; V >= 0; V--) bs = bs_gs_stage(bs, 1 << (R - 1 - V), V == 0 ? cap : HB / 2, last && V == 0);
    return bs;
}
template <int R, int HB, bool LAST, u64 BS, int CAP, class TW, int V = R - 1>
FHE_HD void inv_stages16(u64 (&x)[16], const TW& tw, const LimbParams& P) {
    if constexpr (V >= 0) {
        constexpr int st = 1 << (R - 1 - V);
        constexpr int sh = 4 - R + V;
        constexpr bool last = LAST && (V == 0);
        constexpr int limit = (V == 0) ? CAP : HB / 2;
        constexpr int BM = bs_max(BS);                          // every value is below BM q: the offset of the differences
        static_assert(2 * BM <= HB && CAP <= HB / 2, "inverse stage would overflow 64 bits");
        const u64 q = P.q, nq = 0 - q;
        const u64 bq = (BM % kTQ == 0) ? (u64)(BM / kTQ) * P.tq : (u64)BM * q;
#pragma unroll
        for (int key = 0; key < (1 << sh); key++) {
            Twiddle w;
            if (last) { w.w = P.w1ninv; w.ws = P.w1ninv_s; }
            else w = tw.get(V, key);
#pragma unroll
            for (int j = 0; j < st; j++) {
                const int e = (key << (R - V)) | j;
                const int bsum = bs_get(BS, e) + bs_get(BS, e + st);   // a constant once the loops are unrolled
                const u64 X = x[e], Y = x[e + st];
                FHE_BOUND(X, bs_get(BS, e), q); FHE_BOUND(Y, bs_get(BS, e + st), q); FHE_BOUND((unsigned __int128)X + Y, HB, q);
                u64 S = add3z(X, Y);
                const u64 D = X + bq - Y;
                if (last) S = scale_ninv(S, P.nm, q);
                else if (bsum > limit) S = near60_reduce(S, nq);
                x[e] = S;
                x[e + st] = shoup_mul_lazy3(D, w.w, w.ws, nq);
                FHE_BOUND(x[e], bs_get(bs_gs_stage(BS, st, limit, last), e), q); FHE_BOUND(x[e + st], kTQ, q);
            }
        }
        inv_stages16<R, HB, LAST, bs_gs_stage(BS, st, limit, last), CAP, TW, V - 1>(x, tw, P);
    }
}
// One inverse round (R <= 4 stages) on sixteen values from a uniform entry bound BIN: per-element tracking for the near-2^60 moduli,
// inv_stages otherwise.  inv_round16_out = the largest bound it leaves (kTQ when LAST).
template <int R, int HB, bool NEAR, bool LAST, int BIN, int CAP, class TW>
FHE_HD void inv_round16(u64 (&x)[16], const TW& tw, const LimbParams& P) {
    if constexpr (NEAR) inv_stages16<R, HB, LAST, bs_all(BIN), CAP>(x, tw, P);
    else inv_stages<4, R, HB, NEAR, LAST, BIN>(x, tw, P);
}
FHE_HDC int inv_round16_out(int R, int HB, bool NEAR, bool LAST, int bin, int cap) {
    return NEAR ? bs_max(bs_gs_round(bs_all(bin), R, HB, cap, LAST)) : (LAST ? kTQ : inv_bound_after(bin, R, HB, NEAR));
}
// what may stay lazy at the end of the first, second and third round of a four-round inverse (found by enumeration for 4 x 4 stages)
constexpr int kInvCap1 = 4, kInvCap2 = 4, kInvCap3 = 8;

// bring a value bounded by B*q into [0, q)
template <int HB, bool NEAR, int B>
FHE_HD u64 normalize(u64 x, u64 q) {
    if (NEAR && B > 2) return csub_sign(near60_reduce(x, 0 - q), q);
    if (HB >= 16 && B > 8) x = csub(x, 8 * q);
    if (HB >= 8 && B > 4) x = csub(x, 4 * q);
    if (B > 2) x = csub(x, 2 * q);
    if (B > 1) x = csub(x, q);
    return x;
}

// Canonical form of output e of inv_stages<LE, R, ..., LAST = true>: the last stage leaves a scaled sum (scale_ninv: in (0, 2q)) where
// bit R-1 of e is clear -- one conditional subtraction -- and a lazy product in [0, 3q) where it is set.
template <int HB, bool NEAR, int R>
FHE_HD u64 normalize_last(u64 x, int e, u64 q) {
    if (!((e >> (R - 1)) & 1)) return NEAR ? csub_sign(x, q) : csub(x, q);
    return normalize<HB, NEAR, kTQ>(x, q);
}

// =====================================================================================================
// pass B, forward: one CTA (NT = 2^(LB-4) threads) per tile of NB = 2^LB contiguous elements.
// Split into phases separated by a block barrier; phase functions are what the emulator calls.
//   g    : tile base in global memory (in place)
//   s    : NB u64 of shared memory
//   s12/s3 : this tile's staged twiddle blocks (see TwP1/TwP2/TwP3), copied from the plan's per-tile tables
//   B0   : entry bound (1 + 2*K1 after pass A, 1 when K1 = 0)
// =====================================================================================================
template <int LB, int HB, bool NEAR>
struct TileFwd {
    static constexpr int NB = 1 << LB;
    static constexpr int NT = NB / 16;
    static constexpr int R3 = LB - 8;
    static_assert(LB >= 9 && LB <= 12, "tile pass supports 512..4096 elements");

    // phase 1 is split so that the kernel can issue the global loads before it waits for the staged twiddles
    static FHE_HD void phase1_load(u32 tid, const u64* g, u64 (&x)[16]) {
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = ldg1(g + ((e << (LB - 4)) | tid));
    }
    template <int B0>
    static FHE_HD void phase1_compute(u32 tid, u64 (&x)[16], u64* s, const Twiddle* s12, const LimbParams& P) {
        fwd_stages<4, 4, HB, NEAR, B0>(x, TwP1{s12}, P);
#pragma unroll
        for (int e = 0; e < 16; e++) s[swz((e << (LB - 4)) | tid)] = x[e];
    }
    template <int B0>
    static FHE_HD void phase1(u32 tid, const u64* g, u64* s, const Twiddle* s12, const LimbParams& P) {
        u64 x[16];
        phase1_load(tid, g, x);
        phase1_compute<B0>(tid, x, s, s12, P);
    }
    template <int B0>
    static FHE_HD void phase2(u32 tid, u64* s, const Twiddle* s12, const LimbParams& P) {
        const u32 lo = tid & ((1u << (LB - 8)) - 1), hi = tid >> (LB - 8);
        const u32 base = (hi << (LB - 4)) | lo;
        u64 x[16];
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = s[swz(base | (e << (LB - 8)))];
        fwd_stages<4, 4, HB, NEAR, fwd_bound_after(B0, 4, HB, NEAR)>(x, TwP2{s12, hi}, P);
#pragma unroll
        for (int e = 0; e < 16; e++) s[swz(base | (e << (LB - 8)))] = x[e];
    }
    template <int B0>
    static FHE_HD void phase3(u32 tid, u64* s, const Twiddle* s3, const LimbParams& P) {
        u64 x[16];
        const u64 q = P.q;
        const u32 row = tid << 4;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const u32 a = row | ((k ^ (tid & 7)) << 1);
            ld2(s + a, x[2 * k], x[2 * k + 1]);
        }
        constexpr int B3 = fwd_bound_after(B0, 8, HB, NEAR);
        constexpr int BE = fwd_bound_after(B3, R3, HB, NEAR);
        fwd_stages<4, R3, HB, NEAR, B3>(x, TwP3<NT, R3>{s3, tid}, P);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const u32 a = row | ((k ^ (tid & 7)) << 1);
            st2(s + a, normalize<HB, NEAR, BE>(x[2 * k], q), normalize<HB, NEAR, BE>(x[2 * k + 1], q));
        }
    }
    // coalesced 16-byte copy-out of the swizzled tile
    static FHE_HD void phase4(u32 tid, u64* g, const u64* s) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 c = tid + i * NT;                       // logical 16-byte chunk
            const u32 p = ((c >> 3) << 3) | ((c & 7) ^ ((c >> 3) & 7));
            u64 a, b; ld2(s + 2 * p, a, b); st2(g + 2 * c, a, b);
        }
    }
};

// pass B', inverse: mirror image.  phase1 = coalesced copy-in, phase2 = low strides (16 contiguous per thread),
// phase3 = middle, phase4 = high strides + store.  LAST: K1 == 0, i.e. this tile pass ends the transform.
template <int LB, int HB, bool NEAR>
struct TileInv {
    static constexpr int NB = 1 << LB;
    static constexpr int NT = NB / 16;
    static constexpr int R3 = LB - 8;
    static_assert(LB >= 9 && LB <= 12, "tile pass supports 512..4096 elements");

    static FHE_HD void phase1(u32 tid, const u64* g, u64* s) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 c = tid + i * NT;
            const u32 p = ((c >> 3) << 3) | ((c & 7) ^ ((c >> 3) & 7));
            u64 a, b; ldg2(g + 2 * c, a, b); st2(s + 2 * p, a, b);
        }
    }
    static FHE_HD void phase2(u32 tid, u64* s, const Twiddle* s3, const LimbParams& P) {
        u64 x[16];
        const u32 row = tid << 4;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const u32 a = row | ((k ^ (tid & 7)) << 1);
            ld2(s + a, x[2 * k], x[2 * k + 1]);
        }
        inv_stages<4, R3, HB, NEAR, false, 1>(x, TwP3<NT, R3>{s3, tid}, P);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const u32 a = row | ((k ^ (tid & 7)) << 1);
            st2(s + a, x[2 * k], x[2 * k + 1]);
        }
    }
    static FHE_HD void phase3(u32 tid, u64* s, const Twiddle* s12, const LimbParams& P) {
        const u32 lo = tid & ((1u << (LB - 8)) - 1), hi = tid >> (LB - 8);
        const u32 base = (hi << (LB - 4)) | lo;
        u64 x[16];
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = s[swz(base | (e << (LB - 8)))];
        inv_stages<4, 4, HB, NEAR, false, inv_bound_after(1, R3, HB, NEAR)>(x, TwP2{s12, hi}, P);
#pragma unroll
        for (int e = 0; e < 16; e++) s[swz(base | (e << (LB - 8)))] = x[e];
    }
    // phase 4 is split so that the kernel can release the tile buffer (for the next tile's asynchronous copy-in)
    // as soon as every thread holds its inputs
    static FHE_HD void phase4_load(u32 tid, const u64* s, u64 (&x)[16]) {
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = s[swz((e << (LB - 4)) | tid)];
    }
    template <bool LAST>
    static FHE_HD void phase4_compute(u32 tid, u64 (&x)[16], u64* g, const Twiddle* s12, const LimbParams& P) {
        inv_stages<4, 4, HB, NEAR, LAST, inv_bound_after(1, R3 + 4, HB, NEAR)>(x, TwP1{s12}, P);
        // LAST: fully reduce.  Otherwise leave the lazy bound for pass A' (it starts from out_bound()).
#pragma unroll
        for (int e = 0; e < 16; e++) g[(e << (LB - 4)) | tid] = LAST ? normalize_last<HB, NEAR, 4>(x[e], e, P.q) : x[e];
    }
    template <bool LAST>
    static FHE_HD void phase4(u32 tid, u64* g, const u64* s, const Twiddle* s12, const LimbParams& P) {
        u64 x[16];
        phase4_load(tid, s, x);
        phase4_compute<LAST>(tid, x, g, s12, P);
    }
    static FHE_HDC int out_bound() { return inv_bound_after(1, LB, HB, NEAR); }
};

// =====================================================================================================
// pass A (forward) / A' (inverse): 2^K1 rows x V adjacent columns per thread, registers only.
//   g : limb base, col : first column of this thread, row stride = NB elements (V is 1 or 2)
// =====================================================================================================
template <int K1, int V, int LB, int HB, bool NEAR>
struct RowPass {
    static constexpr int NA = 1 << K1;
    static constexpr int NB = 1 << LB;

    static FHE_HD void forward(u64* gout, const u64* gin, u32 col, const Twiddle* tw, const LimbParams& P) {
        u64 x[V][NA];
#pragma unroll
        for (int r = 0; r < NA; r++) {
            if (V == 2) ldg2(gin + (size_t)r * NB + col, x[0][r], x[V - 1][r]);
            else x[0][r] = ldg1(gin + (size_t)r * NB + col);
        }
#pragma unroll
        for (int c = 0; c < V; c++) fwd_stages<K1, K1, HB, NEAR, 1>(x[c], TwGlobal<0>{tw, 1u}, P);
#pragma unroll
        for (int r = 0; r < NA; r++) {
            if (V == 2) st2(gout + (size_t)r * NB + col, x[0][r], x[V - 1][r]);
            else gout[(size_t)r * NB + col] = x[0][r];
        }
    }
    static FHE_HDC int fwd_out_bound() { return fwd_bound_after(1, K1, HB, NEAR); }

    template <int B0>
    static FHE_HD void inverse(u64* g, u32 col, const Twiddle* tw, const LimbParams& P) {
        u64 x[V][NA];
#pragma unroll
        for (int r = 0; r < NA; r++) {
            if (V == 2) ldg2(g + (size_t)r * NB + col, x[0][r], x[V - 1][r]);
            else x[0][r] = ldg1(g + (size_t)r * NB + col);
        }
#pragma unroll
        for (int c = 0; c < V; c++) inv_stages<K1, K1, HB, NEAR, true, B0>(x[c], TwGlobal<0>{tw, 1u}, P);
#pragma unroll
        for (int r = 0; r < NA; r++) {
            if (V == 2) st2(g + (size_t)r * NB + col, normalize_last<HB, NEAR, K1>(x[0][r], r, P.q), normalize_last<HB, NEAR, K1>(x[V - 1][r], r, P.q));
            else g[(size_t)r * NB + col] = normalize_last<HB, NEAR, K1>(x[0][r], r, P.q);
        }
    }
};

}  // namespace fhe_b200
