// 64-bit modular arithmetic for the B200 RNS-NTT engine.
//
// Replaces the reference's 256-bit path on the per-limb hot loop:
//   fhe::add_mod / sub_mod / mul_mod_montgomery   /root/reference/include/bigint.cuh:27-140
//   fhe::ptx::* carry-chain wrappers              /root/reference/kernels/ptx_bigint.cuh:8-117
//   fhe::ct_butterfly / gs_butterfly              /root/reference/include/ntt.cuh:147-167
// with Shoup (precomputed-quotient) twiddle multiplication, Harvey-style lazy
// butterflies and a 128->64 Barrett reduction.  All moduli are odd primes
// below 2^61.
//
// Every function is __host__ __device__: the kernels are thin __global__
// wrappers around these bodies, and tests/emul/ runs the very same bodies on
// the CPU (thread by thread) so index maps can be validated without a GPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FHE_HD __host__ __device__ __forceinline__
#define FHE_D __device__ __forceinline__
#define FHE_HDC __host__ __device__ constexpr
#else
#define FHE_HD inline
#define FHE_D inline
#define FHE_HDC constexpr
#endif

namespace fhe_b200 {

typedef uint64_t u64;
typedef uint32_t u32;

// (w, floor(w * 2^64 / q)) -- one 128-bit load per twiddle
struct alignas(16) Twiddle { u64 w, ws; };

// per-limb constants, one 64-byte record per modulus
struct alignas(64) LimbParams {
    u64 q;        // modulus
    u64 mu_hi;    // floor(2^128 / q) high word
    u64 mu_lo;    // floor(2^128 / q) low word
    u64 ninv;     // N^-1 mod q
    u64 nm;       // ((q - 1) >> logN) | (logN << 56): N^-1 = -(q-1)/N (mod q), see scale_ninv
    u64 w1ninv;   // inv_tab[1] * N^-1 mod q (last inverse stage, bottom output)
    u64 w1ninv_s; // its Shoup companion
    u64 tq;       // kTQ * q: the offset of the lazy butterfly's difference (from memory, so that the compiler cannot
                  // rematerialise it as a multiply on the binding pipe)
};

FHE_HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

FHE_HD void mul128(u64 a, u64 b, u64& hi, u64& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b; hi = __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    lo = (u64)p; hi = (u64)(p >> 64);
#endif
}

// acc(hi:lo) += a*b ; 128-bit wrap-around accumulate (callers guarantee no overflow)
FHE_HD void mac128(u64& hi, u64& lo, u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    // one carry chain, one asm statement (the CC flag does not survive across statements)
    asm("{\n\t"
        ".reg .u64 pl, ph;\n\t"
        "mul.lo.u64 pl, %2, %3;\n\t"
        "mul.hi.u64 ph, %2, %3;\n\t"
        "add.cc.u64 %0, %0, pl;\n\t"
        "addc.u64 %1, %1, ph;\n\t"
        "}"
        : "+l"(lo), "+l"(hi) : "l"(a), "l"(b));
#else
    unsigned __int128 acc = ((unsigned __int128)hi << 64) | lo;
    acc += (unsigned __int128)a * b;
    lo = (u64)acc; hi = (u64)(acc >> 64);
#endif
}

// (hi:lo) += (bh:bl)
FHE_HD void add128(u64& hi, u64& lo, u64 bh, u64 bl) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u64 %0, %0, %2;\n\taddc.u64 %1, %1, %3;" : "+l"(lo), "+l"(hi) : "l"(bl), "l"(bh));
#else
    unsigned __int128 acc = (((unsigned __int128)hi << 64) | lo) + (((unsigned __int128)bh << 64) | bl);
    lo = (u64)acc; hi = (u64)(acc >> 64);
#endif
}

// 192-bit accumulate: (a2:a1:a0) += (b1:b0)
FHE_HD void add192(u64& a2, u64& a1, u64& a0, u64 b1, u64 b0) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u64 %2, %2, 0;"
        : "+l"(a0), "+l"(a1), "+l"(a2) : "l"(b0), "l"(b1));
#else
    unsigned __int128 lo = ((unsigned __int128)a1 << 64) | a0;
    unsigned __int128 s = lo + (((unsigned __int128)b1 << 64) | b0);
    if (s < lo) a2++;
    a0 = (u64)s; a1 = (u64)(s >> 64);
#endif
}

// A zero the compiler cannot see through (device: a __constant__ word, an immediate operand of IADD3).  ptxas implements a
// two-input 64-bit addition on whichever pipe its model finds idle, and picks the FMA pipe (IMAD.X) -- the one that binds the
// butterflies.  A third addend keeps the addition an IADD3/IADD3.X pair on the ALU pipe: add3z(a, b) = a + b + 0.
#if defined(__CUDACC__)
static __constant__ u64 kOpaqueZero;          // zero-initialised, never written
#endif
FHE_HD u64 add3z(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    return a + b + kOpaqueZero;
#else
    return a + b;
#endif
}

// x - m if x >= m else x
FHE_HD u64 csub(u64 x, u64 m) { return x >= m ? x - m : x; }

FHE_HD u64 add_mod(u64 a, u64 b, u64 q) { return csub(a + b, q); }
FHE_HD u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
FHE_HD u64 neg_mod(u64 a, u64 q) { return a ? q - a : 0; }

// Shoup multiplication, lazy: returns x*w mod q + {0,q}, i.e. a value in [0, 2q).  Valid for ANY 64-bit x.
FHE_HD u64 shoup_mul_lazy(u64 x, u64 w, u64 ws, u64 q) {
    u64 h = mulhi64(x, ws);
    return x * w - h * q;
}
FHE_HD u64 shoup_mul(u64 x, u64 w, u64 ws, u64 q) { return csub(shoup_mul_lazy(x, w, ws, q), q); }

// Shoup multiplication with a three-product quotient estimate: returns x*w mod q + {0,1,2} q, a value in [0, 3q), for ANY
// 64-bit x.  nq = 2^64 - q.
//   h' = x1*s1 + floor((x1*s0 + x0*s1) / 2^32)   (x = x1:x0, ws = s1:s0; the 65-bit cross sum keeps its carry)
// is the true hi64(x*ws) minus {0,1} (only hi32(x0*s0) is dropped), and the true quotient leaves a remainder in [0, 2q), so
// x*w - h'*q is in [0, 3q).  kTQ below is that bound in units of q; every lazy butterfly bound is derived from it.
// On the B200 integer pipe (measured, tools/microbench2.cu): IMAD.lo issues at 64 lanes/clk/SM, IMAD.WIDE and IMAD.HI at 32.
// Written so that every 64-bit temporary is a natural register pair produced by IMAD.WIDE: 5 WIDE + 4 lo = 28 pipe-cycles per
// warp and NOTHING else on the FMA pipe (the earlier form, hi32 products added through a {a,0} pair, cost one register zeroing
// and one move per butterfly there: 30.4 cycles, ncu sm__pipe_fmaheavy_cycles_active).  Carry adds stay on the ALU pipe.
constexpr int kTQ = 3;
FHE_HD u64 shoup_mul_lazy3(u64 x, u64 w, u64 ws, u64 nq) {
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("{\n\t"
        ".reg .u32 y0, y1, s0, s1, w0, w1, n0, n1, c, h0, h1, t0, t1, u0, u1;\n\t"
        ".reg .u64 h, t, U, V;\n\t"
        "mov.b64 {y0, y1}, %1;\n\t"
        "mov.b64 {w0, w1}, %2;\n\t"
        "mov.b64 {s0, s1}, %3;\n\t"
        "mov.b64 {n0, n1}, %4;\n\t"
        "mul.wide.u32 U, y1, s0;\n\t"
        "mul.wide.u32 V, y0, s1;\n\t"
        "mul.wide.u32 h, y1, s1;\n\t"
        "mov.b64 {h0, h1}, h;\n\t"
        "add.cc.u64 U, U, V;\n\t"
        "addc.u32 h1, h1, 0;\n\t"
        "mov.b64 {u0, u1}, U;\n\t"
        "add.cc.u32 h0, h0, u1;\n\t"
        "addc.u32 h1, h1, 0;\n\t"
        "mul.wide.u32 t, y0, w0;\n\t"
        "mad.wide.u32 t, h0, n0, t;\n\t"
        "mov.b64 {t0, t1}, t;\n\t"
        "mad.lo.u32 t1, y0, w1, t1;\n\t"
        "mad.lo.u32 t1, y1, w0, t1;\n\t"
        "mad.lo.u32 t1, h0, n1, t1;\n\t"
        "mad.lo.u32 t1, h1, n0, t1;\n\t"
        "mov.b64 %0, {t0, t1};\n\t"
        "}" : "=l"(r) : "l"(x), "l"(w), "l"(ws), "l"(nq));
    return r;
#else
    const u64 x0 = (u32)x, x1 = x >> 32, s0 = (u32)ws, s1 = ws >> 32;
    const unsigned __int128 cross = (unsigned __int128)(x1 * s0) + (x0 * s1);
    const u64 h = x1 * s1 + (u64)(cross >> 32);
    return x * w + h * nq;
#endif
}

// (For q = 2^60 - delta the partial product h0 * n1 of h * nq is a shift, h0 * 0xF0000000 = -(h0 << 28) mod 2^32, which would leave
// 5 WIDE + 3 lo = 26 cycles of the FMA pipe.  Built and measured in round 2: bit-exact and 4 % SLOWER -- the SHF, LOP3 and IADD3 that
// replace the IMAD cost more issue time than the two pipe cycles saved; see "issue-cost model" in DESIGN.md.)

// (hi:lo) mod q for any 128-bit input; mu = floor(2^128/q), q < 2^63.
// Quotient estimate from the three high partial products; it is low by at most 3, fixed by two conditional
// subtractions (2q then q).  Needs 4q < 2^64.
FHE_HD u64 barrett128(u64 hi, u64 lo, u64 q, u64 mu_hi, u64 mu_lo) {
    // qhat = floor( (hi:lo) * (mu_hi:mu_lo) / 2^128 )  (low 64 bits are all that is needed)
    u64 t1 = mulhi64(lo, mu_hi);
    u64 t2 = mulhi64(hi, mu_lo);
    // dropped: the fractional parts of the two cross terms and lo*mu_lo/2^128 (< 3 in total), and
    // floor(x/q) - floor(x*mu/2^128) <= 1  ==>  0 <= floor(x/q) - qhat <= 3  ==>  r in [0, 4q)
    u64 qhat = hi * mu_hi + t1 + t2;
    u64 r = lo - qhat * q;
    r = csub(r, 2 * q);
    return csub(r, q);
}

// Range reduction for moduli just below 2^60 (delta = 2^60 - q < 2^32, true for the whole 60-bit prime chain):
// x < 16q  ->  (x mod 2^60) + (x >> 60) * delta  =  x - (x >> 60) * q  in [0, 2q).
// Three instructions (SHF, LOP3, IMAD.WIDE: 4 cycles of the FMA pipe) instead of the six of a 64-bit compare-and-subtract.
// nq = 2^64 - q, whose low word is delta.
FHE_HD u64 near60_reduce(u64 x, u64 nq) {
#if defined(__CUDA_ARCH__)
    // Written word by word: the product has no addend and the high word gets a third (opaque zero) addend.  The natural form --
    // IMAD.WIDE with the masked x as its 64-bit addend -- is split by ptxas into WIDE + IADD3 + IMAD.X whenever the two words of
    // the addend do not already sit in an aligned register pair (6 instead of 4 cycles of the binding pipe); a three-input
    // IADD3.X cannot become an IMAD.X, and the extra IADD3 is on the ALU pipe, which has slack.
    u64 r;
    asm("{\n\t"
        ".reg .u32 x0, x1, k, n0, n1, p0, p1;\n\t"
        ".reg .u64 p;\n\t"
        "mov.b64 {x0, x1}, %1;\n\t"
        "mov.b64 {n0, n1}, %2;\n\t"
        "shr.u32 k, x1, 28;\n\t"
        "and.b32 x1, x1, 0x0fffffff;\n\t"
        "mul.wide.u32 p, k, n0;\n\t"
        "mov.b64 {p0, p1}, p;\n\t"
        "add.cc.u32 x0, x0, p0;\n\t"
        "addc.u32 x1, x1, p1;\n\t"
        "add.u32 x1, x1, %3;\n\t"
        "mov.b64 %0, {x0, x1};\n\t"
        "}" : "=l"(r) : "l"(x), "l"(nq), "r"((u32)kOpaqueZero));
    return r;
#else
    return (x & 0x0fffffffffffffffULL) + (x >> 60) * (u64)(u32)nq;
#endif
}

// x - q if x >= q else x, for x < 2^63 + q: the difference's sign decides (one ISETP instead of a 64-bit compare), and the
// subtraction has a third (opaque zero) addend so that it stays on the ALU pipe
FHE_HD u64 csub_sign(u64 x, u64 q) {
#if defined(__CUDA_ARCH__)
    const u64 t = x - q + kOpaqueZero;
    return (long long)t < 0 ? x : t;
#else
    return csub(x, q);
#endif
}

// s * N^-1 mod q (lazy, in (0, 2q)) for any s < 2^64 with s / N < q, N = 2^lg:  m = (q-1)/N is an integer because q = 1 mod 2N,
// and m * N = -1 (mod q), so N^-1 = -m and  s * N^-1 = (s >> lg) - (s mod N) * m.  (s mod N) * m < q.  One IMAD.WIDE + one IMAD
// instead of the nine of a Shoup multiply: the last inverse stage scales N/2 sums with it.
FHE_HD u64 scale_ninv(u64 s, u64 nm, u64 q) {
    const u32 lg = (u32)(nm >> 56);
    const u64 m = nm & 0x00ffffffffffffffULL;
    const u32 lo = (u32)s & ((1u << lg) - 1);                       // lg <= 17
    const u64 prod = (u64)lo * (u32)m + ((u64)(lo * (u32)(m >> 32)) << 32);      // lo * m < q: the high product fits 32 bits
    return (s >> lg) + q - prod;
}

FHE_HD u64 mul_mod(u64 a, u64 b, const LimbParams& P) {
    u64 hi, lo;
    mul128(a, b, hi, lo);
    return barrett128(hi, lo, P.q, P.mu_hi, P.mu_lo);
}

// ---- counter-based generator (specification in DESIGN.md; the CPU checker restates it independently) ----
FHE_HD u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
// Domain separation (so that no two of call seed, batch item, stream, index can alias): the seed is hashed BEFORE the stream offset is
// added, and the seed of batch item b of a call is a hash of (call seed, b) in a domain of its own -- consecutive or G-spaced caller
// seeds, consecutive streams and consecutive batch items all land on unrelated keys.
FHE_HD u64 rng_key(u64 seed, u64 stream) { return mix64(mix64(seed) + 0x9E3779B97F4A7C15ULL * (stream + 1)); }
FHE_HD u64 rng_item_seed(u64 seed, u64 item) { return mix64(mix64(seed ^ 0x6A09E667F3BCC909ULL) + 0x9E3779B97F4A7C15ULL * (item + 1)); }
FHE_HD u64 rng_at(u64 key, u64 idx) { return mix64(key + 0x9E3779B97F4A7C15ULL * (idx + 1)); }

// ---- keyed generator for keys that leave the test bench: ChaCha20 (RFC 8439 block function) ------------------------------------
// The splitmix generator above is a reproducible counter hash keyed by 64 bits -- fine for tests and benchmarks, NOT for real keys.
// With a 256-bit key set on the context (fhe_b200_bfv_set_rng_key) every random word comes from ChaCha20 instead:
//   block counter = idx / 8, nonce = (seed low, seed high, stream), word idx % 8 of the 64-byte block (little-endian pairs);
// the caller's 64-bit seed then only separates calls (a nonce) and need not be secret.  Same samplers, same streams.
struct RngKey { u32 k[8]; u32 on; };
FHE_HD u32 rotl32(u32 x, int n) { return (x << n) | (x >> (32 - n)); }
#define FHE_QR(a, b, c, d) a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); a += b; d ^= a; d = rotl32(d, 8); c += d; b ^= c; b = rotl32(b, 7);
FHE_HD u64 chacha20_word(const RngKey& rk, u32 counter, u32 n0, u32 n1, u32 n2, u32 word) {
    const u32 i0 = 0x61707865u, i1 = 0x3320646eu, i2 = 0x79622d32u, i3 = 0x6b206574u;
    u32 x0 = i0, x1 = i1, x2 = i2, x3 = i3, x4 = rk.k[0], x5 = rk.k[1], x6 = rk.k[2], x7 = rk.k[3], x8 = rk.k[4], x9 = rk.k[5],
        x10 = rk.k[6], x11 = rk.k[7], x12 = counter, x13 = n0, x14 = n1, x15 = n2;
#pragma unroll 1
    for (int r = 0; r < 10; r++) {
        FHE_QR(x0, x4, x8, x12) FHE_QR(x1, x5, x9, x13) FHE_QR(x2, x6, x10, x14) FHE_QR(x3, x7, x11, x15)
        FHE_QR(x0, x5, x10, x15) FHE_QR(x1, x6, x11, x12) FHE_QR(x2, x7, x8, x13) FHE_QR(x3, x4, x9, x14)
    }
    x0 += i0; x1 += i1; x2 += i2; x3 += i3; x4 += rk.k[0]; x5 += rk.k[1]; x6 += rk.k[2]; x7 += rk.k[3];
    x8 += rk.k[4]; x9 += rk.k[5]; x10 += rk.k[6]; x11 += rk.k[7]; x12 += counter; x13 += n0; x14 += n1; x15 += n2;
    u32 lo, hi;
    switch (word & 7u) {
        case 0: lo = x0; hi = x1; break;   case 1: lo = x2; hi = x3; break;   case 2: lo = x4; hi = x5; break;   case 3: lo = x6; hi = x7; break;
        case 4: lo = x8; hi = x9; break;   case 5: lo = x10; hi = x11; break; case 6: lo = x12; hi = x13; break; default: lo = x14; hi = x15; break;
    }
    return ((u64)hi << 32) | lo;
}
#undef FHE_QR
// word idx of stream `stream` of the call / batch item seeded `seed`: the one entry point of every sampler
FHE_HD u64 rng_word(const RngKey& rk, u64 seed, u64 stream, u64 idx) {
    if (!rk.on) return rng_at(rng_key(seed, stream), idx);
    return chacha20_word(rk, (u32)(idx >> 3), (u32)seed, (u32)(seed >> 32), (u32)stream, (u32)idx);
}

}  // namespace fhe_b200
