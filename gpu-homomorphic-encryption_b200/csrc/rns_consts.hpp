// Host precomputation of the constants of an RNS linear combination (exact base conversion, t/Q scale-and-round).
// Fills the placeholders of RNSContext::precompute_crt_parameters (/root/reference/src/rns.cu:42-54: M_i = Q/q_i,
// M_i^-1 mod q_i) and provides the "conversion_matrix" fast_base_conversion_kernel expects
// (/root/reference/include/rns.cuh:116-125).  Only modular arithmetic is needed: no multi-precision integers.
#pragma once
#include <cstdint>
#include <vector>

namespace fhe_b200 {

struct LincombConsts {
    uint32_t S = 0, T = 0;
    bool use_pre = false, use_extra = false;
    std::vector<uint64_t> src_mod, dst_mod;
    std::vector<uint64_t> pre;            // [S]  (Q/q_i)^-1 mod q_i           (conversion only)
    std::vector<uint64_t> th_hi, th_lo;   // [S]  128-bit fixed-point fraction theta_i
    std::vector<uint64_t> M;              // [S][T]
    std::vector<uint64_t> c;              // [T]  multiplier of the rounded integer I
    std::vector<uint64_t> lam;            // [T]  multiplier of the extra limb (scale-and-round only)
};

// exact centred conversion: x (mod prod src) -> x mod dst_k, x taken in [-Q/2, Q/2)
LincombConsts make_conv_consts(const uint64_t* src, uint32_t S, const uint64_t* dst, uint32_t T);
// out_k = round(t/Q * d) mod target_k, d known modulo qs (and ps when with_extra; then targets == ps)
LincombConsts make_scale_consts(const uint64_t* qs, uint32_t L, const uint64_t* ps, uint32_t R, uint64_t t,
                                const uint64_t* targets, uint32_t T, bool with_extra);

}  // namespace fhe_b200
