/*
 * oracle/orc_math.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Plain-C modular arithmetic used by the CPU restatement of the reference's
 * RNS-NTT / BFV path.  Nothing in the product (gpu-homomorphic-encryption_b200/)
 * includes, links or calls this; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do.
 *
 * Follows the reference's *semantics* (not its code): the reference computes
 * (a+b) mod q, (a-b) mod q and a*b mod q on 256-bit words
 * (/root/reference/include/bigint.cuh:27-140); every modulus on the hot path
 * is < 2^61, so the oracle uses uint64_t with unsigned __int128 products.
 */
#ifndef ORC_MATH_H
#define ORC_MATH_H
#include <stdint.h>
#include <stddef.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint32_t u32;

/* (a+b) mod q -- reference add_mod, bigint.cuh:27-50 */
static inline u64 orc_addmod(u64 a, u64 b, u64 q) { u64 s = a + b; return s >= q ? s - q : s; }
/* (a-b) mod q -- reference sub_mod, bigint.cuh:52-73 */
static inline u64 orc_submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
/* a*b mod q -- intent of reference mul_mod_montgomery, bigint.cuh:76-140 (plain residue, no Montgomery scaling) */
static inline u64 orc_mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
static inline u64 orc_negmod(u64 a, u64 q) { return a ? q - a : 0; }

static inline u64 orc_powmod(u64 b, u64 e, u64 q) {
    u64 r = 1 % q; b %= q;
    while (e) { if (e & 1) r = orc_mulmod(r, b, q); b = orc_mulmod(b, b, q); e >>= 1; }
    return r;
}
/* modular inverse for prime q (Fermat) -- intent of reference mod_inverse, src/ntt.cu:110-119 */
static inline u64 orc_invmod_prime(u64 a, u64 q) { return orc_powmod(a % q, q - 2, q); }

static inline u32 orc_bitrev(u32 x, u32 bits) {
    u32 r = 0;
    for (u32 i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
static inline u32 orc_log2(u32 n) { u32 l = 0; while ((1u << l) < n) l++; return l; }

/* floor(num * 2^128 / q) for num < q  -> (hi, lo) 64-bit words */
static inline void orc_frac128(u64 num, u64 q, u64 *hi, u64 *lo) {
    u128 r = ((u128)num << 64);
    u64 h = (u64)(r / q);
    u128 rem = r % q;
    u64 l = (u64)((rem << 64) / q);
    *hi = h; *lo = l;
}

/* splitmix64 finaliser + counter-based generator shared (by specification,
 * not by code) with the CUDA samplers. */
static inline u64 orc_mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* key(seed, stream) = mix64(mix64(seed) + G (stream + 1)): the seed is hashed before the stream offset is added, so G-spaced or
 * consecutive seeds and consecutive streams cannot alias;  word idx = mix64(key + G (idx + 1)). */
static inline u64 orc_rng_key(u64 seed, u64 stream) { return orc_mix64(orc_mix64(seed) + 0x9E3779B97F4A7C15ULL * (stream + 1)); }
/* Keyed mode (test infrastructure mirror of fhe_b200_bfv_set_rng_key): when a 256-bit key is set (orc_set_rng_key), word idx is
 * word idx % 8 of the ChaCha20 block (RFC 8439 block function, 20 rounds) with counter idx / 8 and nonce (seed low, seed high, stream). */
extern u32 orc_chacha_key[8];
extern int orc_chacha_on;
static inline u32 orc_rotl32(u32 x, int n) { return (x << n) | (x >> (32 - n)); }
static inline void orc_chacha20_block(const u32 key[8], u32 counter, const u32 nonce[3], u32 out[16]) {
    u32 st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                  counter, nonce[0], nonce[1], nonce[2]};
    u32 x[16];
    for (int i = 0; i < 16; i++) x[i] = st[i];
#define ORC_QR(a, b, c, d) x[a] += x[b]; x[d] = orc_rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = orc_rotl32(x[b] ^ x[c], 12); \
                           x[a] += x[b]; x[d] = orc_rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = orc_rotl32(x[b] ^ x[c], 7);
    for (int r = 0; r < 10; r++) {
        ORC_QR(0, 4, 8, 12) ORC_QR(1, 5, 9, 13) ORC_QR(2, 6, 10, 14) ORC_QR(3, 7, 11, 15)
        ORC_QR(0, 5, 10, 15) ORC_QR(1, 6, 11, 12) ORC_QR(2, 7, 8, 13) ORC_QR(3, 4, 9, 14)
    }
#undef ORC_QR
    for (int i = 0; i < 16; i++) out[i] = x[i] + st[i];
}
static inline u64 orc_rng64(u64 seed, u64 stream, u64 idx) {
    if (orc_chacha_on) {
        const u32 nonce[3] = {(u32)seed, (u32)(seed >> 32), (u32)stream};
        u32 o[16];
        orc_chacha20_block(orc_chacha_key, (u32)(idx >> 3), nonce, o);
        const u32 w = (u32)idx & 7u;
        return ((u64)o[2 * w + 1] << 32) | o[2 * w];
    }
    return orc_mix64(orc_rng_key(seed, stream) + 0x9E3779B97F4A7C15ULL * (idx + 1));
}
/* seed of batch item `item` of a call seeded with `seed` (a domain of its own) */
static inline u64 orc_item_seed(u64 seed, u64 item) {
    return orc_mix64(orc_mix64(seed ^ 0x6A09E667F3BCC909ULL) + 0x9E3779B97F4A7C15ULL * (item + 1));
}

#endif
