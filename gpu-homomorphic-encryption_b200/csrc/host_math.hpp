// Host-side number theory for plan construction (table and constant precomputation).
// Replaces the reference's placeholders: find_primitive_root returns 3, mod_inverse returns 1,
// twiddles are 1,2,3,... (/root/reference/src/ntt.cu:77-119); compute_montgomery_params
// (/root/reference/src/bigint.cu:23-55); generate_rns_primes / is_prime / find_ntt_prime are
// declared only (/root/reference/include/rns.cuh:139-149).
#pragma once
#include <cstdint>
#include <vector>

namespace fhe_b200 {
namespace host {

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint32_t u32;

inline u64 mulmod(u64 a, u64 b, u64 m) { return (u64)(((u128)a * b) % m); }
inline u64 addmod(u64 a, u64 b, u64 m) { u64 s = a + b; return (s >= m || s < a) ? s - m : s; }
inline u64 submod(u64 a, u64 b, u64 m) { return a >= b ? a - b : a + m - b; }
inline u64 powmod(u64 b, u64 e, u64 m) {
    u64 r = 1 % m; b %= m;
    for (; e; e >>= 1) { if (e & 1) r = mulmod(r, b, m); b = mulmod(b, b, m); }
    return r;
}
// inverse of a modulo m for any coprime pair (extended Euclid); 0 if not invertible
inline u64 invmod(u64 a, u64 m) {
    __int128 r0 = m, r1 = a % m, t0 = 0, t1 = 1;
    while (r1 != 0) { __int128 k = r0 / r1, r2 = r0 - k * r1, t2 = t0 - k * t1; r0 = r1; r1 = r2; t0 = t1; t1 = t2; }
    if (r0 != 1) return 0;
    return (u64)(t0 < 0 ? t0 + m : t0);
}
inline u32 ilog2(u32 n) { u32 l = 0; while ((1u << l) < n) l++; return l; }
inline u32 bitrev(u32 x, u32 bits) { u32 r = 0; for (u32 i = 0; i < bits; i++) { r = (r << 1) | ((x >> i) & 1); } return r; }

// deterministic Miller-Rabin, exact for every 64-bit n (first twelve prime bases)
inline bool is_prime(u64 n) {
    if (n < 2) return false;
    const u64 small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (u64 p : small) { if (n == p) return true; if (n % p == 0) return false; }
    u64 d = n - 1; int s = 0;
    while ((d & 1) == 0) { d >>= 1; s++; }
    for (u64 a : small) {
        u64 x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool witness = true;
        for (int r = 1; r < s && witness; r++) { x = mulmod(x, x, n); if (x == n - 1) witness = false; }
        if (witness) return false;
    }
    return true;
}

// 2N-th primitive root: smallest x >= 2 that is a quadratic non-residue, raised to (q-1)/2N (SURVEY 8d rule)
inline u64 find_psi(u64 q, u32 n) {
    if ((q - 1) % (2ull * n) != 0) return 0;
    for (u64 x = 2; x < q; x++)
        if (powmod(x, (q - 1) / 2, q) == q - 1) return powmod(x, (q - 1) / (2ull * n), q);
    return 0;
}

// descending chain of primes below 2^bits congruent to 1 modulo 2^step_log2
inline std::vector<u64> prime_chain(u32 bits, u32 step_log2, u32 count) {
    std::vector<u64> out;
    const u64 step = 1ull << step_log2;
    for (u64 p = (1ull << bits) - step + 1; out.size() < count && p > step; p -= step)
        if (is_prime(p)) out.push_back(p);
    return out;
}

// floor(w * 2^64 / q), w < q
inline u64 shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }

// floor(num * 2^128 / q) for num < q
inline void frac128(u64 num, u64 q, u64& hi, u64& lo) {
    u128 r = (u128)num << 64;
    hi = (u64)(r / q);
    u128 rem = r % q;
    lo = (u64)((rem << 64) / q);
}

}  // namespace host
}  // namespace fhe_b200
