// Counter-based samplers (device) and their host twins used for key generation.
// Replaces the placeholder sample_gaussian/uniform/ternary kernels (/root/reference/src/polynomial.cu:113-143,
// include/polynomial.cuh:112-135; sample_ternary_kernel is declared only).
//
// Specification (DESIGN.md "Randomness"):
//   r(seed, stream, idx) = mix64(mix64(seed + G*(stream+1)) + G*(idx+1)),  G = 0x9E3779B97F4A7C15, mix64 = splitmix64 finaliser
//   ternary(thr)  : 0 if (r >> 32) >= thr, else (r & 1 ? -1 : +1)                       (thr = P(nonzero) * 2^32)
//   gaussian(cdt) : magnitude = #{k < len-1 : (r >> 1) >= cdt[k]}, sign = r & 1         (cdt[k] = 2^63 P(|X| <= k))
//   uniform(q)    : (r(.., 2 idx) * 2^64 + r(.., 2 idx + 1)) mod q
//   ternary_hw(h) : partial Fisher-Yates over stream (position i swaps with i + r(i) mod (n-i)), signs from stream+1
#pragma once
#include "modarith.cuh"

namespace fhe_b200 {

constexpr int kMaxCdt = 128;

FHE_HD int ternary_from(u64 r, u32 thr) {
    if ((u32)(r >> 32) >= thr) return 0;
    return (r & 1) ? -1 : 1;
}
FHE_HD int gauss_from(u64 r, const u64* cdt, u32 len) {
    const u64 u = r >> 1;
    int mag = 0;
    for (u32 k = 0; k + 1 < len; k++) mag += (u >= cdt[k]) ? 1 : 0;
    return (r & 1) ? -mag : mag;
}
// small signed value -> residue modulo q (|v| < q)
FHE_HD u64 small_to_residue(int v, u64 q) { return v >= 0 ? (u64)v : q - (u64)(-v); }

}  // namespace fhe_b200
