// RNS linear combination (exact base conversion / scale-and-round): object and internal launcher.
#pragma once
#include "common.cuh"
#include "rns_consts.hpp"

struct fhe_b200_lincomb {
    int device = 0;
    uint32_t S = 0, T = 0, SP = 0;          // SP: S padded to a supported register-array size
    bool use_pre = false, use_extra = false;
    fhe_b200::LincombConsts h;              // host copy (constants getter, tests)
    uint64_t* d_blob = nullptr;             // all device constants, one allocation
    // device views into d_blob
    const uint64_t *src_mod, *pre, *pre_s, *th_hi, *th_lo;     // [SP]
    const uint64_t *dst_mod, *mu_hi, *mu_lo, *c, *lam;          // [T]
    const uint64_t* Mt;                                         // [T][SP]
    const uint64_t* Mt30;                                       // [T][SP], entries re-split at bit 30 (SPLIT30 path)
    bool split30 = false;                                       // every modulus below 2^60
    // second copy of the source-side constants padded for the two-lanes-per-coefficient kernel (SP2 = 2 * SPT2)
    uint32_t SPT2 = 0;
    const uint64_t *src_mod2, *pre2, *pre_s2, *th_hi2, *th_lo2, *Mt2, *Mt30_2;
    int force_tpc = 0;                                          // FHE_B200_LINCOMB_TPC = 1 | 2
    const uint32_t *id_src, *id_dst;                            // identity limb maps [SP], [T]
    int sm_count = 148;
    // tensor-core path (lincomb_mma.cu): B fragments of the Toeplitz byte matrix, [ceil(T/4)][mma_kt][8][32] uint2
    uint2* d_bfrag = nullptr;
    uint32_t mma_kt = 0;
    bool use_mma = false;                                       // default when S*T >= 64; FHE_B200_LINCOMB_MMA = 0 | 1 overrides
    // tcgen05 / TMEM path (lincomb_tc.cu): B operand of the Toeplitz byte GEMM in the canonical shared-memory layout
    uint8_t* d_tc_b = nullptr;
    bool use_tc = false;                                        // when the operands fit shared memory; FHE_B200_LINCOMB_TC = 0 | 1 overrides
    // FOLD form of the tcgen05 path (every target modulus in (2^60 - 2^32, 2^60)): eight columns per target carrying the bytes of
    // M[i][k] * 2^(8a) mod m_k, an 82-bit sum folded at 2^60 (lincomb_tc.cu).  [T]: Shoup companion of lam.
    // FHE_B200_LINCOMB_TC_FOLD = 0 keeps the Toeplitz form with the 128-bit Barrett epilogue (the form of every other modulus).
    uint64_t* d_tc_fold = nullptr;
    bool tc_fold = false;
};

namespace fhe_b200 {

// Where the limbs live.  Polynomial b of the input starts at in + b*in_stride; source limb i is limb src_idx[i] of
// it; same for out/extra/sub/add with dst_idx.  Epilogue (ModDown): out = (sub - r) * epi_k [+ add]  (mod m_k).
struct LcView {
    const uint64_t* in = nullptr; size_t in_stride = 0; const uint32_t* src_idx = nullptr;      // device idx arrays
    uint64_t* out = nullptr; size_t out_stride = 0; const uint32_t* dst_idx = nullptr;
    const uint64_t* extra = nullptr; size_t extra_stride = 0;                                    // indexed by dst_idx_extra
    const uint32_t* extra_idx = nullptr;
    const uint64_t* sub = nullptr; size_t sub_stride = 0;                                        // epilogue operands, indexed by epi_idx
    const uint64_t* add = nullptr; size_t add_stride = 0;
    const uint32_t* epi_idx = nullptr;
    const uint64_t* epi_scalar = nullptr;                                                        // [T] device
    const uint64_t* epi_scalar_shoup = nullptr;                                                  // [T] device, floor(epi_scalar * 2^64 / m): optional (cheaper product)
    // pass-through: the raw source limbs are also written to copy_out (limb copy_idx[i] of polynomial b at copy_out + b*copy_stride);
    // saves the separate device-to-device copy when the sources are part of the output polynomial (Q limbs next to the R limbs)
    uint64_t* copy_out = nullptr; size_t copy_stride = 0; const uint32_t* copy_idx = nullptr;
    // scattered destinations (limb-sharded execution, shard.cu): when out_tab is set, target k of polynomial b is written to the
    // ABSOLUTE address out_tab[2k] + ((out_poly0 + b) * out_tab[2k+1] + j) * 8 -- every target limb may live in a different
    // buffer, e.g. the peer GPU that owns that limb (stores over NVLink).  copy_tab does the same for the pass-through limbs
    // ([S] pairs; copy_out must still be non-null to enable the pass-through).  Device arrays of {address, words per polynomial}.
    // a second input added to the first modulo the source moduli before anything else (same limb map as `in`): decryption's c0 + c1 s.
    // IMAD kernel only (lincomb_launch routes such calls there).
    const uint64_t* in_add = nullptr; size_t in_add_stride = 0;
    const uint64_t* out_tab = nullptr; size_t out_poly0 = 0;
    const uint64_t* copy_tab = nullptr; size_t copy_poly0 = 0;
};

int lincomb_create(const LincombConsts& consts, int device, fhe_b200_lincomb** out);
int lincomb_launch(fhe_b200_lincomb* lc, const LcView& v, uint32_t n, uint32_t batch, cudaStream_t st);
// tensor-core path; v must have every default (index maps, strides) filled in
int lincomb_mma_launch(fhe_b200_lincomb* lc, const LcView& v, uint32_t n, uint32_t batch, cudaStream_t st);
// tcgen05 path; n must be a multiple of 128
int lincomb_tc_launch(fhe_b200_lincomb* lc, const LcView& v, uint32_t n, uint32_t batch, cudaStream_t st);
size_t lincomb_tc_smem_bytes(uint32_t S, uint32_t T, bool fold);
void lincomb_tc_build_b(const LincombConsts& h, uint32_t KS, bool fold, std::vector<uint8_t>& out);
uint32_t lincomb_mma_pad_kt(uint32_t S);
void lincomb_mma_build_bfrag(const LincombConsts& h, uint32_t KT, std::vector<uint2>& out);

}  // namespace fhe_b200
