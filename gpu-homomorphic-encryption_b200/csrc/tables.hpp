// Per-limb NTT tables (host construction).  Intent of NTTEngine::precompute_twiddle_factors,
// /root/reference/src/ntt.cu:77-108 (placeholder there): fwd[k] = psi^bitrev(k), inv[k] = psi^-bitrev(k),
// each with its Shoup companion, plus N^-1 folded constants for the last inverse stage.
#pragma once
#include "host_math.hpp"
#include "modarith.cuh"

namespace fhe_b200 {

// returns 0 on success; -1 if q is not an odd prime < 2^61 with q = 1 (mod 2N)
int build_limb_tables(uint64_t q, uint32_t n, Twiddle* fwd, Twiddle* inv, LimbParams* P);
// tile-pass geometry for a ring of 2^logn: the tile pass covers LB = min(logn, 12) stages, the row pass the first K1 = logn - LB
inline int tile_lb(uint32_t logn) { return logn < 12 ? (int)logn : 12; }
// Re-lays the twiddles of the tile pass out per tile so that one bulk copy stages each block into shared memory
// (layouts: ntt_core.cuh, TwP1/TwP2/TwP3).  main: the limb's N-entry table of one direction.
//   p12: [2^K1 tiles][256]     p3: [2^K1 tiles][p3_entries(LB)]
void build_tile_tables(const Twiddle* main, uint32_t logn, Twiddle* p12, Twiddle* p3);
size_t tile_p3_entries(uint32_t logn);
// Balanced two-pass NTT (ntt_bal.cuh), 2^13 <= N <= 2^16: one 512-entry block per pair of adjacent 256-element tiles
// (layout: ntt_bal.cuh, TwB1/TwB2).  out: [N/512 pairs][512]
void build_bal_tables(const Twiddle* main, uint32_t logn, Twiddle* out);
// head-room of the lazy butterflies for a set of moduli: 16 if all < 2^60, else 8 (all < 2^61)
int lazy_headroom(const uint64_t* moduli, uint32_t count);
// true when every modulus lies in (2^60 - 2^32, 2^60) (the whole deterministic 60-bit prime chain does)
bool all_near60(const uint64_t* moduli, uint32_t count);

}  // namespace fhe_b200
