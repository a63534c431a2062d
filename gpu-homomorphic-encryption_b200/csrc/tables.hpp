// Per-limb NTT tables (host construction).  Intent of NTTEngine::precompute_twiddle_factors,
// /root/reference/src/ntt.cu:77-108 (placeholder there): fwd[k] = psi^bitrev(k), inv[k] = psi^-bitrev(k),
// each with its Shoup companion, plus N^-1 folded constants for the last inverse stage.
#pragma once
#include "host_math.hpp"
#include "modarith.cuh"

namespace fhe_b200 {

// returns 0 on success; -1 if q is not an odd prime < 2^61 with q = 1 (mod 2N)
int build_limb_tables(uint64_t q, uint32_t n, Twiddle* fwd, Twiddle* inv, LimbParams* P);
// head-room of the lazy butterflies for a set of moduli: 16 if all < 2^60, else 8 (all < 2^61)
int lazy_headroom(const uint64_t* moduli, uint32_t count);

}  // namespace fhe_b200
