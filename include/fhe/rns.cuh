// Compat header: fhe::RNSContext and the prime utilities (reference: include/rns.cuh:27-65,139-149, src/rns.cu).
// Residue arrays are limb-major [num_primes][count], each residue stored as uint256_t like the reference's API
// (the reference mixes limb-major and coefficient-major indexing; limb-major is the one kept).
// `count` must be a power of two in [512, 131072] and the primes NTT-friendly for it (q = 1 mod 2*count): the
// context is backed by an engine plan for that ring degree.
#pragma once
#include <map>
#include <vector>
#include "ntt.cuh"

namespace fhe {

struct RNSBase {
    uint32_t num_primes = 0;
    std::vector<uint256_t> primes;
};

namespace detail {
inline uint64_t mulmod64(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)(((unsigned __int128)a * b) % m); }
inline uint64_t powmod64(uint64_t b, uint64_t e, uint64_t m) { uint64_t r = 1 % m; b %= m; for (; e; e >>= 1) { if (e & 1) r = mulmod64(r, b, m); b = mulmod64(b, b, m); } return r; }
}  // namespace detail

// deterministic Miller-Rabin for 64-bit values (the `iterations` argument of the reference signature is ignored)
inline bool is_prime(const uint256_t& v, uint32_t = 20) {
    if (!detail::fits_u64(v)) return false;
    const uint64_t n = v.limbs[0];
    if (n < 2) return false;
    const uint64_t bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (uint64_t p : bases) { if (n == p) return true; if (n % p == 0) return false; }
    uint64_t d = n - 1; int s = 0;
    while (!(d & 1)) { d >>= 1; s++; }
    for (uint64_t a : bases) {
        uint64_t x = detail::powmod64(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int r = 1; r < s && comp; r++) { x = detail::mulmod64(x, x, n); if (x == n - 1) comp = false; }
        if (comp) return false;
    }
    return true;
}
// largest prime below 2^bit_length with q = 1 (mod 2 * ntt_size)
inline uint256_t find_ntt_prime(uint32_t bit_length, uint32_t ntt_size) {
    const uint64_t step = 2ull * ntt_size;
    for (uint64_t p = (((1ull << bit_length) - 1) / step) * step + 1; p > step; p -= step)
        if (is_prime(uint256_t(p))) return uint256_t(p);
    throw std::runtime_error("find_ntt_prime: none found");
}
inline std::vector<uint256_t> generate_rns_primes(uint32_t bit_length, uint32_t num_primes, uint32_t ntt_size) {
    std::vector<uint256_t> out;
    const uint64_t step = 2ull * ntt_size;
    for (uint64_t p = (((1ull << bit_length) - 1) / step) * step + 1; p > step && out.size() < num_primes; p -= step)
        if (is_prime(uint256_t(p))) out.push_back(uint256_t(p));
    if (out.size() < num_primes) throw std::runtime_error("generate_rns_primes: not enough primes");
    return out;
}

class RNSContext {
public:
    explicit RNSContext(const std::vector<uint256_t>& primes) {
        base_.num_primes = (uint32_t)primes.size(); base_.primes = primes;
        for (auto& p : primes) { if (!detail::fits_u64(p)) throw std::runtime_error("RNSContext: primes must be below 2^61"); q_.push_back(p.limbs[0]); }
        detail::check_cuda(cudaStreamCreate(&stream_), "cudaStreamCreate");
    }
    ~RNSContext() { for (auto& kv : plans_) fhe_b200_plan_destroy(kv.second); if (stream_) cudaStreamDestroy(stream_); }
    RNSContext(const RNSContext&) = delete;
    RNSContext& operator=(const RNSContext&) = delete;

    void to_rns(uint256_t* d_rns_residues, const uint256_t* d_values, uint32_t count) {
        const size_t tot = (size_t)count * q_.size(); a_.reserve(tot);
        detail::check(fhe_b200_to_rns_u256(plan(count), a_.p, d_values, count, 0, (uint32_t)q_.size(), stream_), "to_rns");
        detail::check(fhe_b200_pack_u256(d_rns_residues, a_.p, tot, stream_), "pack");
    }
    void from_rns(uint256_t* d_values, const uint256_t* d_rns_residues, uint32_t count) {
        const size_t tot = (size_t)count * q_.size(); a_.reserve(tot);
        detail::check(fhe_b200_unpack_u256(a_.p, d_rns_residues, tot, stream_), "unpack");
        detail::check(fhe_b200_from_rns_u256(plan(count), d_values, a_.p, count, 0, (uint32_t)q_.size(), stream_), "from_rns");
    }
    void add_rns(uint256_t* r, const uint256_t* a, const uint256_t* b, uint32_t count) { binary(r, a, b, count, 0); }
    void sub_rns(uint256_t* r, const uint256_t* a, const uint256_t* b, uint32_t count) { binary(r, a, b, count, 1); }
    void mul_rns(uint256_t* r, const uint256_t* a, const uint256_t* b, uint32_t count) { binary(r, a, b, count, 2); }
    // drop limbs [new_level, old_level) one at a time with rounding; input old_level limbs, output new_level limbs
    void mod_switch_rns(uint256_t* d_result, const uint256_t* d_input, uint32_t old_level, uint32_t new_level, uint32_t count) {
        if (new_level == 0 || new_level > old_level || old_level > q_.size()) throw std::runtime_error("mod_switch_rns: bad levels");
        a_.reserve((size_t)count * old_level); b_.reserve((size_t)count * old_level);
        detail::check(fhe_b200_unpack_u256(a_.p, d_input, (size_t)count * old_level, stream_), "unpack");
        uint64_t *src = a_.p, *dst = b_.p;
        for (uint32_t lv = old_level; lv > new_level; lv--) {
            detail::check(fhe_b200_modswitch_drop_last(plan(count), dst, src, 1, 0, lv, stream_), "mod_switch_rns");
            uint64_t* t = src; src = dst; dst = t;
        }
        detail::check(fhe_b200_pack_u256(d_result, src, (size_t)count * new_level, stream_), "pack");
    }
    // exact extension to another basis
    void base_extend(uint256_t* d_extended, const uint256_t* d_input, const RNSBase& target_base, uint32_t count) {
        std::vector<uint64_t> tq;
        for (auto& p : target_base.primes) tq.push_back(p.limbs[0]);
        int dev = 0; cudaGetDevice(&dev);
        fhe_b200_lincomb* lc = nullptr;
        detail::check(fhe_b200_lincomb_create_conv(q_.data(), (uint32_t)q_.size(), tq.data(), (uint32_t)tq.size(), dev, &lc), "base_extend");
        a_.reserve((size_t)count * q_.size()); b_.reserve((size_t)count * tq.size());
        detail::check(fhe_b200_unpack_u256(a_.p, d_input, (size_t)count * q_.size(), stream_), "unpack");
        const int rc = fhe_b200_lincomb_apply(lc, b_.p, a_.p, nullptr, count, 1, stream_);
        if (rc == 0) detail::check(fhe_b200_pack_u256(d_extended, b_.p, (size_t)count * tq.size(), stream_), "pack");
        cudaStreamSynchronize(stream_);
        fhe_b200_lincomb_destroy(lc);
        detail::check(rc, "base_extend");
    }
    uint32_t num_primes() const { return base_.num_primes; }
    const RNSBase& base() const { return base_; }
    cudaStream_t stream() const { return stream_; }

private:
    fhe_b200_plan* plan(uint32_t count) {
        auto it = plans_.find(count);
        if (it != plans_.end()) return it->second;
        int dev = 0; detail::check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
        fhe_b200_plan* p = nullptr;
        detail::check(fhe_b200_plan_create(count, q_.data(), (uint32_t)q_.size(), dev, &p), "RNSContext");
        plans_[count] = p;
        return p;
    }
    void binary(uint256_t* r, const uint256_t* a, const uint256_t* b, uint32_t count, int op) {
        const size_t tot = (size_t)count * q_.size();
        a_.reserve(tot); b_.reserve(tot);
        detail::check(fhe_b200_unpack_u256(a_.p, a, tot, stream_), "unpack");
        detail::check(fhe_b200_unpack_u256(b_.p, b, tot, stream_), "unpack");
        fhe_b200_plan* p = plan(count);
        const uint32_t k = (uint32_t)q_.size();
        detail::check(op == 0 ? fhe_b200_poly_add(p, a_.p, a_.p, b_.p, 1, 0, k, stream_)
                    : op == 1 ? fhe_b200_poly_sub(p, a_.p, a_.p, b_.p, 1, 0, k, stream_)
                              : fhe_b200_poly_mul(p, a_.p, a_.p, b_.p, 1, 0, k, stream_), "rns op");
        detail::check(fhe_b200_pack_u256(r, a_.p, tot, stream_), "pack");
    }
    RNSBase base_;
    std::vector<uint64_t> q_;
    std::map<uint32_t, fhe_b200_plan*> plans_;
    cudaStream_t stream_ = nullptr;
    detail::DeviceBuf a_, b_;
};

}  // namespace fhe
