// Compat header: fhe::Polynomial / fhe::PolynomialOps (reference: include/polynomial.cuh:10-59, src/polynomial.cu).
#pragma once
#include "ntt.cuh"
#include <cmath>
#include <vector>

namespace fhe {

// Coefficient-form polynomial on the device.
//  - created through the reference constructor (degree, modulus < 2^61): `coeffs` (uint256_t[degree], zeroed) is the
//    storage, exactly as in the reference (which allocates degree+1 words, src/polynomial.cu:6-10);
//  - created by FHEContext for an RNS modulus Q = q_0...q_{L-1}: `rns` (uint64_t [limbs][degree], limb-major) is the
//    storage and `coeffs` is null; materialize_coeffs() produces the CRT-composed 256-bit view when Q < 2^256.
struct Polynomial {
    uint256_t* coeffs = nullptr;
    uint32_t degree = 0;
    uint256_t modulus;
    bool is_ntt_form = false;
    uint64_t* rns = nullptr;      // engine-native storage (may be a view into a larger allocation)
    uint32_t limbs = 0;
    bool owns_rns = false;

    Polynomial(uint32_t deg, const uint256_t& mod) : degree(deg), modulus(mod) {
        detail::check_cuda(cudaMalloc(&coeffs, (size_t)deg * sizeof(uint256_t)), "Polynomial");
        detail::check_cuda(cudaMemset(coeffs, 0, (size_t)deg * sizeof(uint256_t)), "Polynomial");
    }
    // RNS polynomial; view != nullptr: non-owning window into a key / ciphertext allocation
    Polynomial(uint32_t deg, uint32_t n_limbs, uint64_t* view = nullptr) : degree(deg), limbs(n_limbs) {
        if (view) { rns = view; owns_rns = false; }
        else {
            detail::check_cuda(cudaMalloc(&rns, (size_t)deg * n_limbs * sizeof(uint64_t)), "Polynomial");
            detail::check_cuda(cudaMemset(rns, 0, (size_t)deg * n_limbs * sizeof(uint64_t)), "Polynomial");
            owns_rns = true;
        }
    }
    ~Polynomial() { if (coeffs) cudaFree(coeffs); if (rns && owns_rns) cudaFree(rns); }
    Polynomial(const Polynomial&) = delete;               // the reference's shallow copies double-free
    Polynomial& operator=(const Polynomial&) = delete;
};

class PolynomialOps {
public:
    PolynomialOps(uint32_t max_degree, const uint256_t& modulus, NTTEngine* ntt) : max_degree_(max_degree), modulus_(modulus), ntt_engine_(ntt) {
        if (!ntt) throw std::runtime_error("PolynomialOps: null NTTEngine");
    }
    void add(Polynomial& result, const Polynomial& a, const Polynomial& b) { binary(result, a, b, 0); }
    void sub(Polynomial& result, const Polynomial& a, const Polynomial& b) { binary(result, a, b, 1); }
    void mul(Polynomial& result, const Polynomial& a, const Polynomial& b) { mul_ntt(result, a, b); }
    void mul_ntt(Polynomial& result, const Polynomial& a, const Polynomial& b) { ntt_engine_->multiply(result.coeffs, a.coeffs, b.coeffs); }
    void mul_negacyclic(Polynomial& result, const Polynomial& a, const Polynomial& b) { mul_ntt(result, a, b); }
    void mul_scalar(Polynomial& result, const Polynomial& a, const uint256_t& scalar) { scalar_op(result, a, scalar, true); }
    void add_scalar(Polynomial& result, const Polynomial& a, const uint256_t& scalar) { scalar_op(result, a, scalar, false); }
    // result = round(q'/q * a) mod q', coefficients of a taken in [0, q)  (declared only in the reference: include/polynomial.cuh:42,
    // poly_mod_switch_kernel :96-102; "in decrypt q' = t").  Exact (integer fixed point); q' < 2^61 coprime to q.
    void mod_switch(Polynomial& result, const Polynomial& a, const uint256_t& new_modulus) {
        const uint32_t n = ntt_engine_->degree();
        const uint64_t q = modulus_.limbs[0], qn = new_modulus.limbs[0];
        if (new_modulus.limbs[1] | new_modulus.limbs[2] | new_modulus.limbs[3]) throw std::runtime_error("mod_switch: new modulus must be below 2^61");
        if (!ms_ || ms_to_ != qn) {
            if (ms_) fhe_b200_lincomb_destroy(ms_);
            ms_ = nullptr;
            int dev = 0; detail::check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
            detail::check(fhe_b200_lincomb_create_scale(&q, 1, nullptr, 0, qn, &qn, 1, 0, dev, &ms_), "mod_switch");
            ms_to_ = qn;
        }
        ua_.reserve(n); ub_.reserve(n);
        cudaStream_t st = ntt_engine_->stream();
        detail::check(fhe_b200_unpack_u256(ua_.p, a.coeffs, n, st), "unpack");
        detail::check(fhe_b200_lincomb_apply(ms_, ub_.p, ua_.p, nullptr, n, 1, st), "mod_switch");
        detail::check(fhe_b200_pack_u256(result.coeffs, ub_.p, n, st), "pack");
        result.modulus = new_modulus;
    }
    // log2 of the largest centred coefficient magnitude (the reference declares `double estimate_noise(const Polynomial&)`,
    // include/polynomial.cuh:45, without defining or using it); host-side convenience, synchronises the stream
    double estimate_noise(const Polynomial& poly) {
        const uint32_t n = ntt_engine_->degree();
        std::vector<uint256_t> h(n);
        detail::check_cuda(cudaStreamSynchronize(ntt_engine_->stream()), "estimate_noise");
        detail::check_cuda(cudaMemcpy(h.data(), poly.coeffs, (size_t)n * sizeof(uint256_t), cudaMemcpyDeviceToHost), "estimate_noise");
        const uint64_t q = modulus_.limbs[0];
        uint64_t worst = 0;
        for (uint32_t i = 0; i < n; i++) { const uint64_t v = h[i].limbs[0] % q; const uint64_t c = v > q / 2 ? q - v : v; if (c > worst) worst = c; }
        return worst ? std::log2((double)worst) : 0.0;
    }
    ~PolynomialOps() { if (ms_) fhe_b200_lincomb_destroy(ms_); }

private:
    void binary(Polynomial& r, const Polynomial& a, const Polynomial& b, int op) {
        const uint32_t n = ntt_engine_->degree();
        ua_.reserve(n); ub_.reserve(n);
        cudaStream_t st = ntt_engine_->stream();
        detail::check(fhe_b200_unpack_u256(ua_.p, a.coeffs, n, st), "unpack");
        detail::check(fhe_b200_unpack_u256(ub_.p, b.coeffs, n, st), "unpack");
        detail::check(op == 0 ? fhe_b200_poly_add(ntt_engine_->plan(), ua_.p, ua_.p, ub_.p, 1, 0, 1, st)
                              : fhe_b200_poly_sub(ntt_engine_->plan(), ua_.p, ua_.p, ub_.p, 1, 0, 1, st), "poly op");
        detail::check(fhe_b200_pack_u256(r.coeffs, ua_.p, n, st), "pack");
    }
    void scalar_op(Polynomial& r, const Polynomial& a, const uint256_t& scalar, bool mul) {
        const uint32_t n = ntt_engine_->degree();
        ua_.reserve(n);
        cudaStream_t st = ntt_engine_->stream();
        // reduce the 256-bit scalar modulo the (64-bit) modulus on the host
        const uint64_t q = modulus_.limbs[0];
        unsigned __int128 acc = 0;
        for (int i = 3; i >= 0; i--) acc = ((acc << 64) | scalar.limbs[i]) % q;
        const uint64_t s = (uint64_t)acc;
        detail::check(fhe_b200_unpack_u256(ua_.p, a.coeffs, n, st), "unpack");
        detail::check(mul ? fhe_b200_poly_mul_scalar(ntt_engine_->plan(), ua_.p, ua_.p, &s, 1, 0, 1, st)
                          : fhe_b200_poly_add_scalar(ntt_engine_->plan(), ua_.p, ua_.p, &s, 1, 0, 1, st), "scalar op");
        detail::check(fhe_b200_pack_u256(r.coeffs, ua_.p, n, st), "pack");
    }
    uint32_t max_degree_;
    uint256_t modulus_;
    NTTEngine* ntt_engine_;
    detail::DeviceBuf ua_, ub_;
    fhe_b200_lincomb* ms_ = nullptr;      // cached mod_switch constants for the last target modulus
    uint64_t ms_to_ = 0;
};

}  // namespace fhe
