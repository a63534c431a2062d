// Compat header: fhe::NTTEngine / fhe::RNS_NTTEngine with the reference's signatures (include/ntt.cuh:72-137),
// implemented on the C ABI of libfhe_b200.so.  Device arrays are uint256_t as in the reference; they are unpacked
// to 8-byte residues at the edge, transformed by the engine and packed back.  Native entry points on uint64_t
// buffers (the fast path; no edge conversion) are the *_u64 methods.
// NTT-domain order: forward() leaves the values in bit-reversed order and inverse() consumes that order (the
// reference's own NTT output is not a transform -- src/ntt.cu:86-97 -- so no caller can depend on its order).
#pragma once
#include <vector>
#include "bigint.cuh"

namespace fhe {

namespace detail {
struct DeviceBuf {
    uint64_t* p = nullptr; size_t words = 0;
    void reserve(size_t w) {
        if (w <= words) return;
        if (p) cudaFree(p);
        check_cuda(cudaMalloc(&p, w * sizeof(uint64_t)), "cudaMalloc");
        words = w;
    }
    ~DeviceBuf() { if (p) cudaFree(p); }
};
}  // namespace detail

class NTTEngine {
public:
    NTTEngine(uint32_t polynomial_degree, const uint256_t& modulus) : n_(polynomial_degree), modulus_(modulus) {
        if (!detail::fits_u64(modulus)) throw std::runtime_error("NTTEngine: modulus must be below 2^61");
        const uint64_t q = modulus.limbs[0];
        int dev = 0; detail::check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
        detail::check(fhe_b200_plan_create(n_, &q, 1, dev, &plan_), "NTTEngine");
        detail::check_cuda(cudaStreamCreate(&stream_), "cudaStreamCreate");
    }
    ~NTTEngine() { fhe_b200_plan_destroy(plan_); if (stream_) cudaStreamDestroy(stream_); }
    NTTEngine(const NTTEngine&) = delete;
    NTTEngine& operator=(const NTTEngine&) = delete;

    void forward(uint256_t* d_data) { forward_batch(d_data, 1); }
    void inverse(uint256_t* d_data) { inverse_batch(d_data, 1); }
    // result = a * b in Z_q[x]/(x^n + 1); result may alias an input (src/ntt.cu:49-75)
    void multiply(uint256_t* d_result, const uint256_t* d_a, const uint256_t* d_b) {
        a_.reserve(n_); b_.reserve(n_);
        detail::check(fhe_b200_unpack_u256(a_.p, d_a, n_, stream_), "unpack");
        detail::check(fhe_b200_unpack_u256(b_.p, d_b, n_, stream_), "unpack");
        detail::check(fhe_b200_negacyclic_mul(plan_, a_.p, a_.p, b_.p, 1, 0, 1, stream_), "multiply");
        detail::check(fhe_b200_pack_u256(d_result, a_.p, n_, stream_), "pack");
    }
    void forward_batch(uint256_t* d_data, uint32_t batch_size) { run(d_data, batch_size, false); }
    void inverse_batch(uint256_t* d_data, uint32_t batch_size) { run(d_data, batch_size, true); }

    // native fast path: uint64_t [batch][n], in place
    void forward_u64(uint64_t* d, uint32_t batch = 1) { detail::check(fhe_b200_ntt_forward(plan_, d, d, batch, 0, 1, stream_), "forward"); }
    void inverse_u64(uint64_t* d, uint32_t batch = 1) { detail::check(fhe_b200_ntt_inverse(plan_, d, d, batch, 0, 1, stream_), "inverse"); }
    void multiply_u64(uint64_t* r, const uint64_t* a, const uint64_t* b, uint32_t batch = 1) {
        detail::check(fhe_b200_negacyclic_mul(plan_, r, a, b, batch, 0, 1, stream_), "multiply");
    }
    cudaStream_t stream() const { return stream_; }
    fhe_b200_plan* plan() const { return plan_; }
    uint32_t degree() const { return n_; }
    const uint256_t& modulus() const { return modulus_; }

private:
    void run(uint256_t* d_data, uint32_t batch, bool inv) {
        const size_t cnt = (size_t)n_ * batch;
        a_.reserve(cnt);
        detail::check(fhe_b200_unpack_u256(a_.p, d_data, cnt, stream_), "unpack");
        detail::check(inv ? fhe_b200_ntt_inverse(plan_, a_.p, a_.p, batch, 0, 1, stream_)
                          : fhe_b200_ntt_forward(plan_, a_.p, a_.p, batch, 0, 1, stream_), "ntt");
        detail::check(fhe_b200_pack_u256(d_data, a_.p, cnt, stream_), "pack");
    }
    uint32_t n_;
    uint256_t modulus_;
    fhe_b200_plan* plan_ = nullptr;
    cudaStream_t stream_ = nullptr;
    detail::DeviceBuf a_, b_;
};

// RNS-based NTT: limb-major d_rns_data[i * n + j] is residue i of coefficient j (src/ntt.cu:158-164)
class RNS_NTTEngine {
public:
    RNS_NTTEngine(uint32_t polynomial_degree, const uint256_t* rns_moduli, uint32_t num_primes) : n_(polynomial_degree), k_(num_primes) {
        std::vector<uint64_t> q(num_primes);
        for (uint32_t i = 0; i < num_primes; i++) {
            if (!detail::fits_u64(rns_moduli[i])) throw std::runtime_error("RNS_NTTEngine: moduli must be below 2^61");
            q[i] = rns_moduli[i].limbs[0];
        }
        int dev = 0; detail::check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
        detail::check(fhe_b200_plan_create(n_, q.data(), k_, dev, &plan_), "RNS_NTTEngine");
        detail::check_cuda(cudaStreamCreate(&stream_), "cudaStreamCreate");
    }
    ~RNS_NTTEngine() { fhe_b200_plan_destroy(plan_); if (stream_) cudaStreamDestroy(stream_); }
    RNS_NTTEngine(const RNS_NTTEngine&) = delete;
    RNS_NTTEngine& operator=(const RNS_NTTEngine&) = delete;

    // d_data: n coefficients (256-bit) -> d_rns_data: k*n residues, each stored as uint256_t like the reference
    void to_rns(uint256_t* d_rns_data, const uint256_t* d_data) {
        a_.reserve((size_t)k_ * n_);
        detail::check(fhe_b200_to_rns_u256(plan_, a_.p, d_data, n_, 0, k_, stream_), "to_rns");
        detail::check(fhe_b200_pack_u256(d_rns_data, a_.p, (size_t)k_ * n_, stream_), "pack");
    }
    void from_rns(uint256_t* d_data, const uint256_t* d_rns_data) {
        a_.reserve((size_t)k_ * n_);
        detail::check(fhe_b200_unpack_u256(a_.p, d_rns_data, (size_t)k_ * n_, stream_), "unpack");
        detail::check(fhe_b200_from_rns_u256(plan_, d_data, a_.p, n_, 0, k_, stream_), "from_rns");
    }
    void forward_rns(uint256_t* d_rns_data) { run(d_rns_data, false); }
    void inverse_rns(uint256_t* d_rns_data) { run(d_rns_data, true); }
    void multiply_rns(uint256_t* d_result, const uint256_t* d_a, const uint256_t* d_b) {
        const size_t cnt = (size_t)k_ * n_;
        a_.reserve(cnt); b_.reserve(cnt);
        detail::check(fhe_b200_unpack_u256(a_.p, d_a, cnt, stream_), "unpack");
        detail::check(fhe_b200_unpack_u256(b_.p, d_b, cnt, stream_), "unpack");
        detail::check(fhe_b200_negacyclic_mul(plan_, a_.p, a_.p, b_.p, 1, 0, k_, stream_), "multiply_rns");
        detail::check(fhe_b200_pack_u256(d_result, a_.p, cnt, stream_), "pack");
    }
    // native fast path: uint64_t [batch][k][n]
    void forward_rns_u64(uint64_t* d, uint32_t batch = 1) { detail::check(fhe_b200_ntt_forward(plan_, d, d, batch, 0, k_, stream_), "forward_rns"); }
    void inverse_rns_u64(uint64_t* d, uint32_t batch = 1) { detail::check(fhe_b200_ntt_inverse(plan_, d, d, batch, 0, k_, stream_), "inverse_rns"); }
    cudaStream_t stream() const { return stream_; }
    fhe_b200_plan* plan() const { return plan_; }

private:
    void run(uint256_t* d, bool inv) {
        const size_t cnt = (size_t)k_ * n_;
        a_.reserve(cnt);
        detail::check(fhe_b200_unpack_u256(a_.p, d, cnt, stream_), "unpack");
        detail::check(inv ? fhe_b200_ntt_inverse(plan_, a_.p, a_.p, 1, 0, k_, stream_) : fhe_b200_ntt_forward(plan_, a_.p, a_.p, 1, 0, k_, stream_), "ntt");
        detail::check(fhe_b200_pack_u256(d, a_.p, cnt, stream_), "pack");
    }
    uint32_t n_, k_;
    fhe_b200_plan* plan_ = nullptr;
    cudaStream_t stream_ = nullptr;
    detail::DeviceBuf a_, b_;
};

}  // namespace fhe
