// Compat header: fhe::uint256_t and friends, source-compatible with the reference's include/bigint.cuh.
// In this engine uint256_t is a BOUNDARY type only: device arrays in the public API carry 32-byte coefficients
// (/root/reference/include/bigint.cuh:9-24); every hot loop runs on 8-byte residues inside libfhe_b200.so.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <cuda_runtime.h>
#include "../fhe_b200.h"

#if defined(__CUDACC__)
#define FHE_COMPAT_HD __host__ __device__
#else
#define FHE_COMPAT_HD
#endif

namespace fhe {

struct uint256_t {
    uint64_t limbs[4];  // little-endian: limbs[0] is least significant
    FHE_COMPAT_HD uint256_t() { limbs[0] = limbs[1] = limbs[2] = limbs[3] = 0; }
    FHE_COMPAT_HD uint256_t(uint64_t v) { limbs[0] = v; limbs[1] = limbs[2] = limbs[3] = 0; }
    FHE_COMPAT_HD uint256_t(uint64_t l0, uint64_t l1, uint64_t l2, uint64_t l3) { limbs[0] = l0; limbs[1] = l1; limbs[2] = l2; limbs[3] = l3; }
};

namespace detail {
// the reference documents std::runtime_error on failure (docs/API_REFERENCE.md:567-572) but never throws; we do
inline void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + fhe_b200_last_error());
}
inline void check_cuda(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
inline bool fits_u64(const uint256_t& v) { return v.limbs[1] == 0 && v.limbs[2] == 0 && v.limbs[3] == 0; }
FHE_COMPAT_HD inline int cmp256(const uint256_t& a, const uint256_t& b) {
    for (int i = 3; i >= 0; i--) { if (a.limbs[i] < b.limbs[i]) return -1; if (a.limbs[i] > b.limbs[i]) return 1; }
    return 0;
}
FHE_COMPAT_HD inline uint64_t add256(uint256_t& r, const uint256_t& a, const uint256_t& b) {
    uint64_t c = 0;
    for (int i = 0; i < 4; i++) { uint64_t s = a.limbs[i] + c; uint64_t c1 = s < c; s += b.limbs[i]; c = c1 | (s < b.limbs[i]); r.limbs[i] = s; }
    return c;
}
FHE_COMPAT_HD inline uint64_t sub256(uint256_t& r, const uint256_t& a, const uint256_t& b) {
    uint64_t br = 0;
    for (int i = 0; i < 4; i++) { uint64_t d = a.limbs[i] - br; uint64_t b1 = a.limbs[i] < br; uint64_t e = d - b.limbs[i]; br = b1 | (d < b.limbs[i]); r.limbs[i] = e; }
    return br;
}
}  // namespace detail

// (a + b) mod m and (a - b) mod m on full 256-bit words (reference: include/bigint.cuh:27-73; unlike the reference the
// borrow test covers all four limbs)
FHE_COMPAT_HD inline uint256_t add_mod(const uint256_t& a, const uint256_t& b, const uint256_t& modulus) {
    uint256_t r; const uint64_t carry = detail::add256(r, a, b);
    if (carry || detail::cmp256(r, modulus) >= 0) { uint256_t t; detail::sub256(t, r, modulus); return t; }
    return r;
}
FHE_COMPAT_HD inline uint256_t sub_mod(const uint256_t& a, const uint256_t& b, const uint256_t& modulus) {
    uint256_t r; const uint64_t borrow = detail::sub256(r, a, b);
    if (borrow) { uint256_t t; detail::add256(t, r, modulus); return t; }
    return r;
}

struct MontgomeryParams {
    uint256_t modulus;
    uint256_t r_squared;  // R^2 mod N, R = 2^64 for the single-word moduli this engine uses
    uint256_t inv;        // -N^-1 mod 2^64
};
// reference: src/bigint.cu:23-55.  Kept for source compatibility; the engine itself uses Shoup/Barrett, not Montgomery.
inline uint256_t compute_montgomery_inverse(const uint256_t& modulus) {
    uint64_t q = modulus.limbs[0], x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - q * x;        // Newton: x = q^-1 mod 2^64
    return uint256_t(0 - x);
}
inline MontgomeryParams compute_montgomery_params(const uint256_t& modulus) {
    MontgomeryParams p; p.modulus = modulus; p.inv = compute_montgomery_inverse(modulus);
    if (detail::fits_u64(modulus) && modulus.limbs[0] > 1) {
        const unsigned __int128 r = (((unsigned __int128)1) << 64) % modulus.limbs[0];
        p.r_squared = uint256_t((uint64_t)((r * r) % modulus.limbs[0]));
    }
    return p;
}

}  // namespace fhe
