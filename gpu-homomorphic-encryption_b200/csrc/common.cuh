// Shared host-side plumbing of libfhe_b200.so: error reporting, the plan object, launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../include/fhe_b200.h"
#include "modarith.cuh"

namespace fhe_b200 {

// Function attributes (dynamic shared memory limits) are per device: a process may hold plans on several GPUs, so "set once" is
// tracked per device ordinal.
// Every entry point runs on the device its plan / context / conversion object was created for, whatever the caller's current
// device is, and leaves the caller's current device as it found it.
struct DeviceGuard {
    int prev = -1; bool changed = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) changed = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

struct PerDeviceOnce {
    unsigned long long mask = 0;
    bool need(int device) { const unsigned long long bit = 1ull << (device & 63); if (mask & bit) return false; mask |= bit; return true; }
};


void set_error(const char* fmt, ...);

#define FHE_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            fhe_b200::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return FHE_B200_ECUDA;                                                                  \
        }                                                                                           \
    } while (0)

#define FHE_REQUIRE(cond, ...)                                  \
    do {                                                        \
        if (!(cond)) { fhe_b200::set_error(__VA_ARGS__); return FHE_B200_EINVAL; } \
    } while (0)

#define FHE_TRY(expr)                       \
    do { int rc__ = (expr); if (rc__ != 0) return rc__; } while (0)

void count_launch();
// checks the launch that has just been issued (and counts it)
#define FHE_LAUNCH_CHECK() do { fhe_b200::count_launch(); FHE_CUDA(cudaGetLastError()); } while (0)

// optional per-kernel event timing (fhe_b200_profile_enable)
bool profile_on();
void profile_begin(int kind, uint64_t units, cudaStream_t st);
void profile_end(cudaStream_t st);

}  // namespace fhe_b200

// The plan: ring degree, moduli, device tables.  (C struct name = the opaque handle of the C ABI.)
struct fhe_b200_plan {
    uint32_t n = 0, logn = 0, limbs = 0;
    int device = 0;
    int hb = 16;                                   // lazy head-room (16: all q < 2^60, 8: all q < 2^61)
    bool near60 = false;                           // every q in (2^60 - 2^32, 2^60): cheap range reduction (near60_reduce)
    std::vector<uint64_t> moduli;
    fhe_b200::Twiddle* d_fwd = nullptr;            // [limbs][n]
    fhe_b200::Twiddle* d_inv = nullptr;            // [limbs][n]
    // per-tile staged blocks of the tile pass (one bulk copy each): [limbs][tiles][256] and [limbs][tiles][p3]
    fhe_b200::Twiddle *d_fwd_p12 = nullptr, *d_fwd_p3 = nullptr, *d_inv_p12 = nullptr, *d_inv_p3 = nullptr;
    // balanced two-pass NTT (ntt_bal.cuh), 2^13 <= N <= 2^16: per tile-pair staged blocks [limbs][n/512][512]; default path there
    bool bal = false;                              // FHE_B200_NTT_BAL=0 falls back to the row+tile passes
    fhe_b200::Twiddle *d_fwd_bal = nullptr, *d_inv_bal = nullptr;
    size_t p3_entries = 0;                         // per tile
    uint32_t tiles = 1;                            // tiles per limb = 2^K1
    fhe_b200::LimbParams* d_params = nullptr;      // [limbs]
    std::vector<fhe_b200::LimbParams> h_params;
    size_t chunk_bytes = (size_t)1 << 30;          // polynomials per row-pass/tile-pass launch pair (FHE_B200_NTT_CHUNK_MB)
    int sm_count = 148;
    // host-buffer pipeline (lazy)
    cudaStream_t hs[3] = {nullptr, nullptr, nullptr};
    uint64_t* d_stage[3] = {nullptr, nullptr, nullptr};
    size_t stage_bytes = 0;
};

namespace fhe_b200 {
// kernels/launchers implemented in ntt_kernels.cu / elementwise.cu
int launch_ntt(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t batch, uint32_t limb_begin,
               uint32_t limb_count, bool inverse, cudaStream_t st);
enum EwOp { EW_ADD = 0, EW_SUB, EW_MUL, EW_MAC, EW_MUL_SCALAR, EW_ADD_SCALAR, EW_NEG };
int launch_elementwise(fhe_b200_plan* plan, EwOp op, uint64_t* d_out, const uint64_t* d_a, const uint64_t* d_b,
                       const uint64_t* d_c, uint32_t batch, uint32_t limb_begin, uint32_t limb_count, cudaStream_t st);
// Scatter form of the last inverse pass (limb-sharded execution): the coefficients of a limb are cut into 2^k blocks of `nc`
// consecutive coefficients; block h of local limb l of polynomial p is written to
//   base[h] + ((p * limbs_total + limb_off + l) * nc + (coefficient - h * nc)) * 8      (absolute addresses, possibly peer memory)
struct BalScatter {
    uint64_t base[16];
    uint32_t log_blocks;            // log2 of the number of coefficient blocks (<= 4)
    uint32_t limbs_total, limb_off, nc;
};
int launch_ntt_bal(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t limb_begin, uint32_t limb_count,
                   uint32_t l0, uint32_t nl, uint32_t b0, uint32_t nb, bool inverse, cudaStream_t st,
                   const BalScatter* scatter = nullptr);
int launch_ntt_strided_in(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, size_t in_poly_stride, uint32_t batch,
                          uint32_t limb_begin, uint32_t limb_count, bool inverse, cudaStream_t st);
// inverse transform of [batch][limb_count][N] whose tile pass runs in place on d_buf and whose column pass scatters the result
// Forward transforms of the balanced two-pass sizes launched inside this scope leave their results lazy, in [0, 8q) instead of [0, q)
// (ntt_bal.cuh, BalB::fwd_phase2): only for consumers that reduce 128-bit products of them anyway.  Other sizes are unaffected.
extern thread_local bool g_fwd_lazy_out;
struct ScopedLazyForward {
    bool prev;
    explicit ScopedLazyForward(bool on) : prev(g_fwd_lazy_out) { g_fwd_lazy_out = on; }
    ~ScopedLazyForward() { g_fwd_lazy_out = prev; }
};
int launch_ntt_inverse_scatter(fhe_b200_plan* plan, uint64_t* d_buf, uint32_t batch, uint32_t limb_begin, uint32_t limb_count,
                               const BalScatter& scatter, cudaStream_t st);
// Fused tile passes of the BFV multiply (ntt_fused.cu): forward tile pass of the operands, pointwise work, inverse tile pass of the
// results in one kernel.  The operands must have been through the forward COLUMN pass (launch_ntt_pass_a), the results still need
// the inverse column pass (launch_ntt_pass_a with inverse = true, or the scatter form).  Element offsets: operand p of ciphertext b,
// buffer limb l (plan limb limb_begin + l) at in + in_plane[p] + b * in_poly + l * N.
struct FusedTile {
    uint64_t* out = nullptr; const uint64_t* in = nullptr;
    uint32_t limb_begin = 0, limb_count = 0, nb = 0;
    size_t in_plane[4] = {0, 0, 0, 0}, in_poly = 0, out_plane[3] = {0, 0, 0}, out_poly = 0;
    const uint64_t* key = nullptr; size_t key_poly = 0; uint32_t dnum = 0;     // inner product: key polynomial (d, c) at key + (2d + c) * key_poly + l * N
    bool square = false;                                                       // tensor product of a ciphertext with itself: operands 0, 1 only
};
bool fused_tile_supported(const fhe_b200_plan* plan, uint32_t dnum);
int launch_fused_tile(fhe_b200_plan* plan, int mode, const FusedTile& t, cudaStream_t st);
// one pass of the balanced transform on its own: the column pass (first log N - 8 stages forward; last inverse) of [batch][limb_count][N]
int launch_ntt_pass_a(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t batch, uint32_t limb_begin, uint32_t limb_count,
                      bool inverse, cudaStream_t st, const BalScatter* scatter = nullptr);
int launch_negacyclic_mul_fused(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_a, const uint64_t* d_b, uint32_t batch,
                                uint32_t limb_begin, uint32_t limb_count, cudaStream_t st);
int check_range(const fhe_b200_plan* plan, uint32_t batch, uint32_t limb_begin, uint32_t limb_count);
}  // namespace fhe_b200
