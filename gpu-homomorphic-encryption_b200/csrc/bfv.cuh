// The BFV context object behind the opaque C handle (shared by bfv.cu and shard.cu).
#pragma once
#include "common.cuh"
#include "lincomb.cuh"

struct fhe_b200_bfv {
    uint32_t n = 0, logn = 0, L = 0, R = 0, K = 0, dnum = 0, alpha = 0;
    uint64_t t = 0;
    int device = 0;
    uint32_t hw = 0, thr = 1u << 31;
    fhe_b200::RngKey rng = {{0, 0, 0, 0, 0, 0, 0, 0}, 0};   // on = 0: reproducible splitmix generator; fhe_b200_bfv_set_rng_key switches to ChaCha20
    std::vector<uint64_t> primes;
    fhe_b200_plan* plan = nullptr;                       // all L+R primes
    fhe_b200_lincomb *q2r = nullptr, *scale = nullptr, *r2q = nullptr, *moddown = nullptr, *dec = nullptr;
    std::vector<fhe_b200_lincomb*> modup;                // [dnum]
    // device constants
    uint64_t* d_consts = nullptr;                        // delta[L] | p_mod_q[L+K] (0 outside Q) | pinv_mod_q[L] | cdt[128] | Shoup companions of pinv[L]
    const uint64_t *d_delta = nullptr, *d_pmodq = nullptr, *d_pinv = nullptr, *d_pinv_s = nullptr, *d_cdt = nullptr;
    uint32_t cdt_len = 0;
    std::vector<uint64_t> h_cdt;
    uint32_t* d_idx = nullptr;                           // index maps for the lincomb views
    const uint32_t *d_idx_p = nullptr;                   // [K]  L .. L+K-1
    std::vector<const uint32_t*> d_idx_grp, d_idx_tgt;   // per digit: sources [alpha], targets [L+K-alpha]
    // workspaces, grown on demand and kept (a context is single-threaded by contract, see fhe_b200.h)
    uint64_t* d_ws = nullptr; size_t ws_words = 0;       // multiply / encrypt / decrypt scratch
    uint64_t* d_io[3] = {nullptr, nullptr, nullptr}; size_t io_words[3] = {0, 0, 0};   // device staging of the host-buffer entry point
    cudaStream_t io_stream[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t io_done[3] = {nullptr, nullptr, nullptr};
    // multiply: the batch is split in two halves on two internal streams (the conversions of one half overlap the transforms of the other)
    cudaStream_t mul_stream[2] = {nullptr, nullptr};
    cudaEvent_t mul_fork = nullptr, mul_join[2] = {nullptr, nullptr};
    cudaEvent_t mul_ev[3] = {nullptr, nullptr, nullptr};            // single-ciphertext multiply: branch synchronisation
};

namespace fhe_b200 {
int ensure_words(uint64_t** buf, size_t* have, size_t need);
}  // namespace fhe_b200
