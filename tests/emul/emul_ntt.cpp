// TEST INFRASTRUCTURE: runs the engine's kernel bodies (csrc/ntt_core.cuh) on the CPU, one emulated thread at a
// time with phase boundaries where the kernels have block barriers, so that index maps, twiddle addressing and
// lazy-reduction bounds can be validated against the oracle in the GPU-less container.  Not a product path:
// nothing in the C-ABI library references this file.
#include <vector>
#include <cstdlib>
#include <cstring>
#include "ntt_core.cuh"
#include "ntt_bal.cuh"
#include "tables.hpp"

using namespace fhe_b200;

struct TileTabs { std::vector<Twiddle> p12, p3; size_t p3n; };
static TileTabs tile_tabs(const Twiddle* main, u32 logn) {
    TileTabs t;
    const u32 tiles = 1u << (logn - (u32)tile_lb(logn));
    t.p3n = tile_p3_entries(logn);
    t.p12.resize((size_t)tiles * 256); t.p3.resize((size_t)tiles * t.p3n);
    build_tile_tables(main, logn, t.p12.data(), t.p3.data());
    return t;
}

template <int LB, int K1, int HB, bool NEAR>
static void fwd_limb(u64* d, const Twiddle* tw, const LimbParams& P) {
    constexpr int NB = 1 << LB;
    if constexpr (K1 > 0) {
        constexpr int V = (K1 >= 5) ? 1 : 2;
        for (u32 col = 0; col < (u32)NB; col += V) RowPass<K1, V, LB, HB, NEAR>::forward(d, d, col, tw, P);
    }
    constexpr int B0 = fwd_bound_after(1, K1, HB, NEAR);
    const TileTabs tt = tile_tabs(tw, LB + K1);
    std::vector<u64> s(NB);
    using T = TileFwd<LB, HB, NEAR>;
    for (u32 b = 0; b < (1u << K1); b++) {
        u64* g = d + (size_t)b * NB;
        // what the kernel's two bulk copies stage into shared memory
        std::vector<Twiddle> s12(tt.p12.begin() + (size_t)b * 256, tt.p12.begin() + (size_t)(b + 1) * 256);
        std::vector<Twiddle> s3(tt.p3.begin() + (size_t)b * tt.p3n, tt.p3.begin() + (size_t)(b + 1) * tt.p3n);
        for (u32 t = 0; t < (u32)T::NT; t++) T::template phase1<B0>(t, g, s.data(), s12.data(), P);
        for (u32 t = 0; t < (u32)T::NT; t++) T::template phase2<B0>(t, s.data(), s12.data(), P);
        for (u32 t = 0; t < (u32)T::NT; t++) T::template phase3<B0>(t, s.data(), s3.data(), P);
        for (u32 t = 0; t < (u32)T::NT; t++) T::phase4(t, g, s.data());
    }
}

template <int LB, int K1, int HB, bool NEAR>
static void inv_limb(u64* d, const Twiddle* tw, const LimbParams& P) {
    constexpr int NB = 1 << LB;
    const TileTabs tt = tile_tabs(tw, LB + K1);
    std::vector<u64> s(NB);
    using T = TileInv<LB, HB, NEAR>;
    for (u32 b = 0; b < (1u << K1); b++) {
        u64* g = d + (size_t)b * NB;
        std::vector<Twiddle> s12(tt.p12.begin() + (size_t)b * 256, tt.p12.begin() + (size_t)(b + 1) * 256);
        std::vector<Twiddle> s3(tt.p3.begin() + (size_t)b * tt.p3n, tt.p3.begin() + (size_t)(b + 1) * tt.p3n);
        for (u32 t = 0; t < (u32)T::NT; t++) T::phase1(t, g, s.data());
        for (u32 t = 0; t < (u32)T::NT; t++) T::phase2(t, s.data(), s3.data(), P);
        for (u32 t = 0; t < (u32)T::NT; t++) T::phase3(t, s.data(), s12.data(), P);
        for (u32 t = 0; t < (u32)T::NT; t++) T::template phase4<K1 == 0>(t, g, s.data(), s12.data(), P);
    }
    if constexpr (K1 > 0) {
        constexpr int V = (K1 >= 5) ? 1 : 2;
        constexpr int B0 = T::out_bound();
        for (u32 col = 0; col < (u32)NB; col += V) RowPass<K1, V, LB, HB, NEAR>::template inverse<B0>(d, col, tw, P);
    }
}

template <int HB, bool NEAR>
static int run(u64* d, u32 logn, const Twiddle* tw, const LimbParams& P, int inverse) {
#define CASE(LBv, K1v) if (inverse) inv_limb<LBv, K1v, HB, NEAR>(d, tw, P); else fwd_limb<LBv, K1v, HB, NEAR>(d, tw, P); return 0;
    switch (logn) {
        case 9: CASE(9, 0)
        case 10: CASE(10, 0)
        case 11: CASE(11, 0)
        case 12: CASE(12, 0)
        case 13: CASE(12, 1)
        case 14: CASE(12, 2)
        case 15: CASE(12, 3)
        case 16: CASE(12, 4)
        case 17: CASE(12, 5)
    }
#undef CASE
    return -2;
}

extern "C" int emul_ntt(uint64_t* data, uint32_t n, uint64_t q, int inverse, int hb) {
    std::vector<Twiddle> fwd(n), inv(n);
    LimbParams P;
    if (build_limb_tables(q, n, fwd.data(), inv.data(), &P)) return -1;
    const u32 logn = host::ilog2(n);
    const Twiddle* tw = inverse ? inv.data() : fwd.data();
    if (hb == 16 && all_near60(&q, 1)) return run<16, true>(data, logn, tw, P, inverse);     // what the plan would pick
    if (hb == 16) return run<16, false>(data, logn, tw, P, inverse);
    if (hb == 8) return run<8, false>(data, logn, tw, P, inverse);
    return -3;
}

// ---- balanced two-pass bodies (ntt_bal.cuh): pass A per item with a barrier between the rounds, pass B per warp ----
template <int KA, int HB, bool NEAR>
static void bal_limb(u64* d, const Twiddle* tw, const LimbParams& P, int inverse) {
    using A = BalA<KA, HB, NEAR>;
    using B = BalB<HB, NEAR>;
    const u32 logn = KA + 8, pairs = 1u << (KA - 1);
    std::vector<Twiddle> blocks((size_t)pairs * 512);
    build_bal_tables(tw, logn, blocks.data());
    std::vector<u64> s(4096), sw(512);
    std::vector<Twiddle> stw(tw, tw + (1u << KA));          // what the CTA stages
    if (!inverse) {
        for (u32 cb = 0; cb < (u32)A::CB; cb++) {
            u64* g = d + (size_t)cb * A::C;
            for (u32 t = 0; t < 256; t++) A::fwd_round1(t, g, s.data(), stw.data(), P);
            for (u32 t = 0; t < 256; t++) A::fwd_round2(t, g, s.data(), stw.data(), P);
        }
        constexpr int B0 = A::fwd_out_bound();
        for (u32 p = 0; p < pairs; p++) {
            u64* g = d + (size_t)p * 512;
            const Twiddle* sb = blocks.data() + (size_t)p * 512;
            // the kernel's bulk copy lands the tile pair in a staging buffer in natural order; round 1 reads its values from there
            std::vector<u64> stg(g, g + 512);
            for (u32 l = 0; l < 32; l++) {
                u64 x[16], y[16];
                B::fwd_load_staged(l, stg.data(), x);
                B::fwd_load(l, g, y);                                   // (the direct global loads of the fused kernels: same values)
                for (int e = 0; e < 16; e++) if (x[e] != y[e]) std::abort();
                B::template fwd_phase1<B0>(l, x, sw.data(), sb, P);
            }
            for (u32 l = 0; l < 32; l++) B::template fwd_phase2<B0>(l, sw.data(), sb, P);
            for (u32 l = 0; l < 32; l++) B::fwd_phase3(l, g, sw.data());
        }
    } else {
        for (u32 p = 0; p < pairs; p++) {
            u64* g = d + (size_t)p * 512;
            const Twiddle* sb = blocks.data() + (size_t)p * 512;
            {   // staged re-layout (what the kernel does after its bulk copy) against the direct one
                std::vector<u64> stg(g, g + 512), sw2(512);
                for (u32 l = 0; l < 32; l++) B::inv_phase1_staged(l, stg.data(), sw.data());
                for (u32 l = 0; l < 32; l++) B::inv_phase1(l, g, sw2.data());
                if (sw2 != sw) std::abort();
            }
            for (u32 l = 0; l < 32; l++) B::inv_phase2(l, sw.data(), sb, P);
            for (u32 l = 0; l < 32; l++) B::inv_phase3(l, g, sw.data(), sb, P);
        }
        constexpr int BIN = B::inv_out_bound();
        for (u32 cb = 0; cb < (u32)A::CB; cb++) {
            u64* g = d + (size_t)cb * A::C;
            for (u32 t = 0; t < 256; t++) A::template inv_round2<BIN>(t, g, s.data(), stw.data(), P);
            for (u32 t = 0; t < 256; t++) A::template inv_round1<BIN>(t, g, s.data(), stw.data(), P);
        }
    }
}
template <int HB, bool NEAR>
static int run_bal(u64* d, u32 logn, const Twiddle* tw, const LimbParams& P, int inverse) {
    switch (logn) {
        case 13: bal_limb<5, HB, NEAR>(d, tw, P, inverse); return 0;
        case 14: bal_limb<6, HB, NEAR>(d, tw, P, inverse); return 0;
        case 15: bal_limb<7, HB, NEAR>(d, tw, P, inverse); return 0;
        case 16: bal_limb<8, HB, NEAR>(d, tw, P, inverse); return 0;
    }
    return -2;
}
extern "C" int emul_ntt_bal(uint64_t* data, uint32_t n, uint64_t q, int inverse, int hb) {
    std::vector<Twiddle> fwd(n), inv(n);
    LimbParams P;
    if (build_limb_tables(q, n, fwd.data(), inv.data(), &P)) return -1;
    const u32 logn = host::ilog2(n);
    const Twiddle* tw = inverse ? inv.data() : fwd.data();
    if (hb == 16 && all_near60(&q, 1)) return run_bal<16, true>(data, logn, tw, P, inverse);
    if (hb == 16) return run_bal<16, false>(data, logn, tw, P, inverse);
    if (hb == 8) return run_bal<8, false>(data, logn, tw, P, inverse);
    return -3;
}

extern "C" uint64_t emul_mul_mod(uint64_t a, uint64_t b, uint64_t q) {
    LimbParams P; P.q = q; host::frac128(1, q, P.mu_hi, P.mu_lo);
    return mul_mod(a, b, P);
}
extern "C" uint64_t emul_barrett128(uint64_t hi, uint64_t lo, uint64_t q) {
    u64 mh, ml; host::frac128(1, q, mh, ml);
    return barrett128(hi, lo, q, mh, ml);
}
