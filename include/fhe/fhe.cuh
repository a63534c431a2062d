// Compat header: fhe::FHEContext and its PODs with the reference's names and signatures (include/fhe.cuh:15-166),
// implemented on the BFV entry points of libfhe_b200.so.
//
// What changes behind the same API: log_q is honoured (L = ceil(log_q / 60) RNS limbs of
// the deterministic 60-bit prime chain instead of the hard-coded q = 2^60), multiply() applies the t/q scaling, and
// relinearize() performs hybrid key switching instead of dropping c2.  Polynomials of keys and ciphertexts are RNS
// polynomials (Polynomial::rns, limb-major uint64_t [L][N]); plaintexts keep the reference's uint256_t coefficients.
#pragma once
#include <random>
#include <cstdlib>
#include <algorithm>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "bigint.cuh"
#include "ntt.cuh"
#include "polynomial.cuh"
#include "rns.cuh"

namespace fhe {

struct SecurityParams {
    uint32_t lambda;          // security level (recorded; parameter validation is the caller's job, as in the reference)
    uint32_t poly_degree;     // power of two in [512, 131072]
    uint32_t log_q;           // bits of the ciphertext modulus -> ceil(log_q / 60) limbs
    float sigma;              // Gaussian noise standard deviation
    uint32_t hamming_weight;  // non-zero coefficients of the ternary secret (0: each coefficient non-zero with prob. 1/2)
};

struct SchemeParams {
    SecurityParams security;
    uint32_t n;
    uint256_t q;              // Q when it fits 256 bits, else 0 (see rns_moduli)
    uint256_t t;
    uint256_t delta;          // floor(Q/t) when it fits 256 bits, else 0
    RNSContext* rns_ctx = nullptr;
    NTTEngine* ntt_engine = nullptr;      // single-limb engine over q_0 (kept for callers that reach for it)
    PolynomialOps* poly_ops = nullptr;
    uint32_t num_levels = 0;
    std::vector<uint256_t> modulus_chain; // q_0 .. q_{L-1}
    // engine view
    uint32_t L = 0, R = 0, K = 0, dnum = 0;
    std::vector<uint64_t> rns_moduli;     // Q limbs then auxiliary limbs
};

struct PublicKey { Polynomial* pk0 = nullptr; Polynomial* pk1 = nullptr; };
struct SecretKey { Polynomial* sk = nullptr; };
struct RelinKeys { std::vector<PublicKey*> rlk_keys; uint32_t decomp_bits = 0; };
struct GaloisKeys { std::vector<PublicKey*> gal_keys; std::vector<uint32_t> elts; };   // elts[i]: the Galois element of gal_keys[i]
struct Ciphertext {
    std::vector<Polynomial*> components;
    uint32_t level = 0;
    float noise_budget = 0;
    bool is_ntt_form = false;
};
struct Plaintext { Polynomial* poly = nullptr; bool is_ntt_form = false; };

class FHEContext {
public:
    explicit FHEContext(const SecurityParams& sp) {
        params_.security = sp;
        params_.n = sp.poly_degree;
        const uint32_t L = std::max<uint32_t>(1, (sp.log_q + 59) / 60);
        // hybrid key switching: dnum digits of alpha limbs, special modulus of K = alpha auxiliary limbs
        uint32_t dnum = L;
        if (L >= 6 && L % 3 == 0) dnum = 3; else if (L >= 4 && L % 2 == 0) dnum = 2;
        const uint32_t alpha = L / dnum, R = L + 1, K = alpha;
        params_.L = L; params_.R = R; params_.K = K; params_.dnum = dnum;
        // deterministic chain: k-th largest prime below 2^60 with p = 1 (mod 2^18)
        for (uint64_t p = (1ull << 60) - (1ull << 18) + 1; params_.rns_moduli.size() < L + R; p -= (1ull << 18))
            if (is_prime(uint256_t(p))) params_.rns_moduli.push_back(p);
        for (uint32_t i = 0; i < L; i++) params_.modulus_chain.push_back(uint256_t(params_.rns_moduli[i]));
        params_.num_levels = L;
        params_.t = uint256_t(65537);
        cudaGetDevice(&device_);
        detail::check(fhe_b200_bfv_create(params_.n, L, R, K, dnum, 65537, params_.rns_moduli.data(), sp.sigma, sp.hamming_weight,
                                          device_, &ctx_), "FHEContext");
        if (L <= 4) {      // Q and delta as 256-bit numbers for callers that read params().q / params().delta
            uint256_t Q(1);
            for (uint32_t i = 0; i < L; i++) Q = mul_small(Q, params_.rns_moduli[i]);
            params_.q = Q; params_.delta = div_small(Q, 65537);
        }
        std::vector<uint256_t> qs(params_.modulus_chain);
        params_.rns_ctx = new RNSContext(qs);
        params_.ntt_engine = new NTTEngine(params_.n, params_.modulus_chain[0]);
        params_.poly_ops = new PolynomialOps(params_.n, params_.modulus_chain[0], params_.ntt_engine);
        detail::check_cuda(cudaStreamCreate(&stream_), "cudaStreamCreate");
        // randomness: by default a 256-bit ChaCha20 key from the OS (fhe_b200_bfv_set_rng_key) and a fresh 64-bit nonce per call (the
        // secret key's is its own draw); init_rng(seed) or FHE_B200_DETERMINISTIC_SEED switch to the reproducible splitmix generator
        // the tests and the CPU oracle share (DESIGN.md "Randomness").
        if (const char* e = std::getenv("FHE_B200_DETERMINISTIC_SEED")) init_rng(std::strtoull(e, nullptr, 0));
        else {                  // production default: ChaCha20 keyed with 256 bits from the OS
            std::random_device rd; uint8_t key[32];
            for (int i = 0; i < 32; i += 4) { const uint32_t w = rd(); key[i] = (uint8_t)w; key[i + 1] = (uint8_t)(w >> 8); key[i + 2] = (uint8_t)(w >> 16); key[i + 3] = (uint8_t)(w >> 24); }
            detail::check(fhe_b200_bfv_set_rng_key(ctx_, key), "set_rng_key");
        }
    }
    ~FHEContext() {
        delete params_.rns_ctx; delete params_.poly_ops; delete params_.ntt_engine;
        fhe_b200_bfv_destroy(ctx_);
        if (stream_) cudaStreamDestroy(stream_);
    }
    FHEContext(const FHEContext&) = delete;
    FHEContext& operator=(const FHEContext&) = delete;

    // ---- key generation (src/fhe.cu:54-111).  Ownership: the structs own what they receive; call release() helpers.
    void keygen(PublicKey& pk, SecretKey& sk) {
        const uint32_t n = params_.n, L = params_.L, A = params_.L + params_.R;
        sk.sk = new Polynomial(n, A); sk.sk->is_ntt_form = true;
        pk.pk0 = new Polynomial(n, 2 * L); pk.pk0->is_ntt_form = true;                 // owns [2][L][N]
        pk.pk1 = new Polynomial(n, L, pk.pk0->rns + (size_t)L * n); pk.pk1->is_ntt_form = true;
        detail::check(fhe_b200_bfv_keygen(ctx_, next_seed(), next_seed(), sk.sk->rns, pk.pk0->rns, stream_), "keygen");
    }
    // decomp_bits is accepted for source compatibility; the digits are RNS digits (dnum of them), not base-2^w digits
    void relinkey_gen(RelinKeys& rlk, const SecretKey& sk, uint32_t decomp_bits = 16) {
        const uint32_t n = params_.n, W = params_.L + params_.K, dnum = params_.dnum;
        rlk.decomp_bits = decomp_bits;
        Polynomial* all = new Polynomial(n, 2 * W * dnum); all->is_ntt_form = true;    // [dnum][2][W][N]
        detail::check(fhe_b200_bfv_relinkeygen(ctx_, next_seed(), sk.sk->rns, all->rns, stream_), "relinkey_gen");
        for (uint32_t d = 0; d < dnum; d++) {
            PublicKey* k = new PublicKey();
            k->pk0 = d == 0 ? all : new Polynomial(n, W, all->rns + (size_t)(2 * d) * W * n);
            k->pk1 = new Polynomial(n, W, all->rns + (size_t)(2 * d + 1) * W * n);
            rlk.rlk_keys.push_back(k);
        }
    }

    // Galois keys for the power-of-two row rotations (elements 3^(2^k) mod 2N) and the column swap (2N - 1): declared only in
    // the reference (include/fhe.cuh:59-61,86).  Each key has the relinearisation-key layout [dnum][2][L+K][N].
    void galoiskey_gen(GaloisKeys& gal_keys, const SecretKey& sk) {
        const uint32_t n = params_.n, W = params_.L + params_.K, dnum = params_.dnum;
        std::vector<uint32_t> elts;
        uint64_t e = 3;
        for (uint32_t k = 1; k < n / 2; k <<= 1) { elts.push_back((uint32_t)e); e = e * e % (2ull * n); }
        elts.push_back(2 * n - 1);
        for (uint32_t g : elts) {
            PublicKey* k = new PublicKey();
            k->pk0 = new Polynomial(n, 2 * W * dnum); k->pk0->is_ntt_form = true;          // owns [dnum][2][W][N]
            k->pk1 = new Polynomial(n, W, k->pk0->rns + (size_t)W * n); k->pk1->is_ntt_form = true;
            detail::check(fhe_b200_bfv_galoiskeygen(ctx_, next_seed(), g, sk.sk->rns, k->pk0->rns, stream_), "galoiskey_gen");
            gal_keys.gal_keys.push_back(k); gal_keys.elts.push_back(g);
        }
    }
    // slot rotation by `steps` inside each row of the 2 x N/2 batching matrix = automorphism x -> x^(3^steps); composed from the
    // power-of-two keys (include/fhe.cuh:113-114, declared only)
    void rotate_rows(Ciphertext& result, const Ciphertext& ct, int steps, const GaloisKeys& gal_keys) {
        const uint32_t half = params_.n / 2;
        uint32_t s = (uint32_t)(((steps % (int)half) + (int)half) % (int)half);
        Ciphertext cur; new_ciphertext(cur);
        detail::check_cuda(cudaMemcpyAsync(cur.components[0]->rns, ct.components[0]->rns, (size_t)2 * params_.L * params_.n * sizeof(uint64_t),
                                           cudaMemcpyDeviceToDevice, stream_), "rotate_rows");
        for (uint32_t k = 0; (1u << k) < half; k++) {
            if (!((s >> k) & 1u)) continue;
            Ciphertext nxt; new_ciphertext(nxt);
            detail::check(fhe_b200_bfv_apply_galois(ctx_, cur.components[0]->rns, gal_keys.elts.at(k), gal_keys.gal_keys.at(k)->pk0->rns,
                                                    nxt.components[0]->rns, 1, stream_), "rotate_rows");
            detail::check_cuda(cudaStreamSynchronize(stream_), "rotate_rows");
            release(cur); cur = nxt;
        }
        cur.noise_budget = ct.noise_budget; cur.level = ct.level;
        replace(result, cur);
    }
    // swaps the two rows of the batching matrix: x -> x^(2N-1) (include/fhe.cuh:115-116, declared only)
    void rotate_columns(Ciphertext& result, const Ciphertext& ct, const GaloisKeys& gal_keys) {
        Ciphertext out; new_ciphertext(out);
        detail::check(fhe_b200_bfv_apply_galois(ctx_, ct.components[0]->rns, gal_keys.elts.back(), gal_keys.gal_keys.back()->pk0->rns,
                                                out.components[0]->rns, 1, stream_), "rotate_columns");
        out.noise_budget = ct.noise_budget; out.level = ct.level;
        replace(result, out);
    }

    // ---- encoding (src/fhe.cu:113-136): coefficient encoding, values[i] -> coefficient i
    void encode(Plaintext& pt, const std::vector<uint64_t>& values) {
        pt.poly = new Polynomial(params_.n, params_.t);
        pt.is_ntt_form = false;
        std::vector<uint256_t> h(params_.n);
        for (size_t i = 0; i < std::min(values.size(), (size_t)params_.n); i++) h[i] = uint256_t(values[i] % 65537);
        detail::check_cuda(cudaMemcpy(pt.poly->coeffs, h.data(), params_.n * sizeof(uint256_t), cudaMemcpyHostToDevice), "encode");
    }
    void decode(std::vector<uint64_t>& values, const Plaintext& pt) {
        std::vector<uint256_t> h(params_.n);
        detail::check_cuda(cudaStreamSynchronize(stream_), "decode");
        detail::check_cuda(cudaMemcpy(h.data(), pt.poly->coeffs, params_.n * sizeof(uint256_t), cudaMemcpyDeviceToHost), "decode");
        values.clear();
        for (uint32_t i = 0; i < params_.n; i++) values.push_back(h[i].limbs[0]);
    }

    // ---- encryption (src/fhe.cu:138-185)
    void encrypt(Ciphertext& ct, const Plaintext& pt, const PublicKey& pk) {
        const uint32_t n = params_.n, L = params_.L;
        new_ciphertext(ct);
        scratch_.reserve(n);
        detail::check(fhe_b200_unpack_u256(scratch_.p, pt.poly->coeffs, n, stream_), "encrypt");
        detail::check(fhe_b200_bfv_encrypt(ctx_, next_seed(), scratch_.p, pk.pk0->rns, ct.components[0]->rns, 1, stream_), "encrypt");
        ct.level = 0; ct.is_ntt_form = false;
        ct.noise_budget = params_.security.sigma * 10;   // the reference's rough bookkeeping (src/fhe.cu:168)
        (void)L;
    }
    void decrypt(Plaintext& pt, const Ciphertext& ct, const SecretKey& sk) {
        const uint32_t n = params_.n;
        pt.poly = new Polynomial(n, params_.t);
        pt.is_ntt_form = false;
        scratch_.reserve(n);
        detail::check(fhe_b200_bfv_decrypt(ctx_, ct.components[0]->rns, sk.sk->rns, scratch_.p, 1, stream_), "decrypt");
        detail::check(fhe_b200_pack_u256(pt.poly->coeffs, scratch_.p, n, stream_), "decrypt");
    }

    // ---- homomorphic operations (src/fhe.cu:187-235)
    void add(Ciphertext& result, const Ciphertext& a, const Ciphertext& b) {
        if (a.components.size() != b.components.size()) throw std::runtime_error("add: operands differ in their number of components");
        const uint32_t comps = (uint32_t)a.components.size();
        Ciphertext out; new_ciphertext(out, comps);
        if (comps == 2) detail::check(fhe_b200_bfv_add(ctx_, a.components[0]->rns, b.components[0]->rns, out.components[0]->rns, 1, stream_), "add");
        else detail::check(fhe_b200_poly_add(fhe_b200_bfv_plan(ctx_), out.components[0]->rns, a.components[0]->rns, b.components[0]->rns,
                                             comps, 0, params_.L, stream_), "add");
        out.noise_budget = std::min(a.noise_budget, b.noise_budget);
        out.level = std::max(a.level, b.level);
        replace(result, out);
    }
    void multiply(Ciphertext& result, const Ciphertext& a, const Ciphertext& b, const RelinKeys& rlk) {
        Ciphertext out; new_ciphertext(out);
        detail::check(fhe_b200_bfv_multiply_relin(ctx_, a.components[0]->rns, b.components[0]->rns, rlk.rlk_keys.at(0)->pk0->rns,
                                                  out.components[0]->rns, nullptr, 1, stream_), "multiply");
        out.noise_budget = a.noise_budget + b.noise_budget + 10;
        out.level = std::max(a.level, b.level);
        replace(result, out);
    }
    // declared-only in the reference (include/fhe.cuh:98-104): sub, add_plain, sub_plain, multiply_plain
    void sub(Ciphertext& result, const Ciphertext& a, const Ciphertext& b) {
        Ciphertext out; new_ciphertext(out);
        detail::check(fhe_b200_bfv_sub(ctx_, a.components[0]->rns, b.components[0]->rns, out.components[0]->rns, 1, stream_), "sub");
        out.noise_budget = std::min(a.noise_budget, b.noise_budget); out.level = std::max(a.level, b.level);
        replace(result, out);
    }
    void add_plain(Ciphertext& result, const Ciphertext& ct, const Plaintext& pt) { plain_op(result, ct, pt, 0); }
    void sub_plain(Ciphertext& result, const Ciphertext& ct, const Plaintext& pt) { plain_op(result, ct, pt, 1); }
    void multiply_plain(Ciphertext& result, const Ciphertext& ct, const Plaintext& pt) { plain_op(result, ct, pt, 2); }

    // The two halves of the reference's multiply on their own (src/fhe.cu:198-219 builds three components, :226-235 reduces them):
    // multiply without keys returns a 3-component ciphertext; add() accepts two of those; relinearize() brings one back to two
    // components by hybrid key switching.  Lets a sum of products be relinearised once.  &a == &b (or the same buffers) squares.
    void multiply(Ciphertext& result, const Ciphertext& a, const Ciphertext& b) {
        if (a.components.size() != 2 || b.components.size() != 2) throw std::runtime_error("multiply: operands must have two components");
        Ciphertext out; new_ciphertext(out, 3);
        detail::check(fhe_b200_bfv_multiply(ctx_, a.components[0]->rns, b.components[0]->rns, out.components[0]->rns, 1, stream_), "multiply");
        out.noise_budget = a.noise_budget + b.noise_budget + 10;
        out.level = std::max(a.level, b.level);
        replace(result, out);
    }
    void relinearize(Ciphertext& ct, const RelinKeys& rlk) {
        if (ct.components.size() <= 2) return;                                      // src/fhe.cu:227
        Ciphertext out; new_ciphertext(out);
        detail::check(fhe_b200_bfv_relinearize(ctx_, ct.components[0]->rns, rlk.rlk_keys.at(0)->pk0->rns, out.components[0]->rns, 1, stream_), "relinearize");
        out.noise_budget = ct.noise_budget; out.level = ct.level;
        replace(ct, out);
    }

    // invariant noise budget in bits (include/fhe.cuh:142, declared only in the reference)
    float estimate_noise_budget(const Ciphertext& ct, const SecretKey& sk) {
        if (ct.components.size() != 2) throw std::runtime_error("estimate_noise_budget: relinearize the ciphertext first");
        double bits = 0;
        detail::check(fhe_b200_bfv_noise_budget(ctx_, ct.components[0]->rns, sk.sk->rns, 1, &bits, stream_), "estimate_noise_budget");
        return (float)bits;
    }

    const SchemeParams& params() const { return params_; }
    fhe_b200_bfv* engine() const { return ctx_; }
    cudaStream_t stream() const { return stream_; }
    void synchronize() { detail::check_cuda(cudaStreamSynchronize(stream_), "synchronize"); }

    // frees what keygen / encrypt / ... allocated (the reference leaks these)
    static void release(Ciphertext& ct) { for (size_t i = 0; i < ct.components.size(); i++) delete ct.components[i]; ct.components.clear(); }
    static void release(PublicKey& pk) { delete pk.pk1; delete pk.pk0; pk.pk0 = pk.pk1 = nullptr; }
    static void release(SecretKey& sk) { delete sk.sk; sk.sk = nullptr; }
    static void release(Plaintext& pt) { delete pt.poly; pt.poly = nullptr; }
    static void release(GaloisKeys& gk) {
        for (size_t i = gk.gal_keys.size(); i-- > 0;) { delete gk.gal_keys[i]->pk1; delete gk.gal_keys[i]->pk0; delete gk.gal_keys[i]; }
        gk.gal_keys.clear(); gk.elts.clear();
    }
    static void release(RelinKeys& rlk) {
        for (size_t d = rlk.rlk_keys.size(); d-- > 0;) { delete rlk.rlk_keys[d]->pk1; delete rlk.rlk_keys[d]->pk0; delete rlk.rlk_keys[d]; }
        rlk.rlk_keys.clear();
    }

private:
    void plain_op(Ciphertext& result, const Ciphertext& ct, const Plaintext& pt, int op) {
        Ciphertext out; new_ciphertext(out);
        scratch_.reserve(params_.n);
        detail::check(fhe_b200_unpack_u256(scratch_.p, pt.poly->coeffs, params_.n, stream_), "plain operand");
        detail::check(op == 2 ? fhe_b200_bfv_multiply_plain(ctx_, ct.components[0]->rns, scratch_.p, out.components[0]->rns, 1, stream_)
                              : fhe_b200_bfv_add_plain(ctx_, ct.components[0]->rns, scratch_.p, out.components[0]->rns, 1, op, stream_),
                      "plain operand");
        out.noise_budget = ct.noise_budget; out.level = ct.level;
        replace(result, out);
    }
    void new_ciphertext(Ciphertext& ct, uint32_t comps = 2) {
        const uint32_t n = params_.n, L = params_.L;
        Polynomial* c0 = new Polynomial(n, comps * L);                           // owns [comps][L][N]
        ct.components.clear(); ct.components.push_back(c0);
        for (uint32_t k = 1; k < comps; k++) ct.components.push_back(new Polynomial(n, L, c0->rns + (size_t)k * L * n));
    }
    void replace(Ciphertext& dst, Ciphertext& src) {
        // dst may alias an operand: the engine call has been queued on stream_, release after it completes
        if (!dst.components.empty()) { synchronize(); release(dst); }
        dst = src; src.components.clear();
    }
    void init_rng(uint64_t seed) { seed_ = seed; calls_ = 0; deterministic_ = true; fhe_b200_bfv_set_rng_key(ctx_, nullptr); }
    // deterministic mode: a hashed call counter -- call k of this context gets mix(base seed, k); the library hashes the seed again
    // per stream / batch item.  Default mode: 64 fresh bits from the OS per call.
    uint64_t next_seed() {
        if (!deterministic_) { std::random_device rd; return ((uint64_t)rd() << 32) | (uint64_t)rd(); }
        uint64_t z = seed_ + 0x9E3779B97F4A7C15ull * ++calls_;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    static uint256_t mul_small(const uint256_t& a, uint64_t m) {
        uint256_t r; unsigned __int128 c = 0;
        for (int i = 0; i < 4; i++) { c += (unsigned __int128)a.limbs[i] * m; r.limbs[i] = (uint64_t)c; c >>= 64; }
        return r;
    }
    static uint256_t div_small(const uint256_t& a, uint64_t d) {
        uint256_t r; unsigned __int128 rem = 0;
        for (int i = 3; i >= 0; i--) { rem = (rem << 64) | a.limbs[i]; r.limbs[i] = (uint64_t)(rem / d); rem %= d; }
        return r;
    }
    SchemeParams params_;
    fhe_b200_bfv* ctx_ = nullptr;
    cudaStream_t stream_ = nullptr;
    int device_ = 0;
    uint64_t seed_ = 0, calls_ = 0;
    bool deterministic_ = false;
    detail::DeviceBuf scratch_;
};

// SIMD slot encoding (reference: include/fhe.cuh:151-166; src/fhe.cu:267-279 is a stub that forwards to coefficient
// encoding, which is why the reference's printed slot-wise expectations never come out).  Slot i is the value of the
// plaintext polynomial at the evaluation point the engine's negacyclic NTT modulo t leaves at position i, so ciphertext
// add / multiply act slot-wise.  Needs t = 1 (mod 2N): t = 65537 supports N <= 32768.
class BatchEncoder {
public:
    explicit BatchEncoder(const FHEContext& context) : context_(context), slot_count_(context.params().n / 2),
                                                        ntt_(context.params().n, context.params().t) {}
    // up to n values are accepted (slot_count() keeps the reference's n/2)
    void encode(Plaintext& pt, const std::vector<uint64_t>& values) {
        const uint32_t n = context_.params().n;
        std::vector<uint64_t> h(n, 0);
        for (size_t i = 0; i < std::min(values.size(), (size_t)n); i++) h[i] = values[i] % 65537;
        buf_.reserve(n);
        detail::check_cuda(cudaMemcpyAsync(buf_.p, h.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, ntt_.stream()), "encode");
        ntt_.inverse_u64(buf_.p);
        pt.poly = new Polynomial(n, context_.params().t);
        pt.is_ntt_form = false;
        detail::check(fhe_b200_pack_u256(pt.poly->coeffs, buf_.p, n, ntt_.stream()), "encode");
        detail::check_cuda(cudaStreamSynchronize(ntt_.stream()), "encode");
    }
    void decode(std::vector<uint64_t>& values, const Plaintext& pt) {
        const uint32_t n = context_.params().n;
        buf_.reserve(n);
        detail::check_cuda(cudaStreamSynchronize(context_.stream()), "decode");
        detail::check(fhe_b200_unpack_u256(buf_.p, pt.poly->coeffs, n, ntt_.stream()), "decode");
        ntt_.forward_u64(buf_.p);
        values.assign(n, 0);
        detail::check_cuda(cudaMemcpyAsync(values.data(), buf_.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ntt_.stream()), "decode");
        detail::check_cuda(cudaStreamSynchronize(ntt_.stream()), "decode");
    }
    uint32_t slot_count() const { return slot_count_; }

private:
    const FHEContext& context_;
    uint32_t slot_count_;
    NTTEngine ntt_;
    detail::DeviceBuf buf_;
};

// Performance monitoring (include/fhe.cuh:169-198; declared only in the reference).  start_timer / stop_timer bracket work on a
// stream with CUDA events (the context's stream when one is given, else the legacy default stream); get_stats synchronises the
// recorded events and fills the reference's PerfStats fields for the operation names "encrypt", "decrypt", "add", "multiply",
// "relinearize", "mod_switch", "bootstrap"; any other name is timed and listed by print_stats.
struct PerfStats {
    double encrypt_time_ms = 0, decrypt_time_ms = 0, add_time_ms = 0, mul_time_ms = 0, relin_time_ms = 0, mod_switch_time_ms = 0,
           bootstrap_time_ms = 0;
    uint64_t num_encryptions = 0, num_multiplications = 0, num_bootstraps = 0;
};

class PerformanceMonitor {
public:
    explicit PerformanceMonitor(cudaStream_t stream = nullptr) : stream_(stream) {}
    ~PerformanceMonitor() { reset(); }
    PerformanceMonitor(const PerformanceMonitor&) = delete;
    PerformanceMonitor& operator=(const PerformanceMonitor&) = delete;
    void start_timer(const std::string& operation) {
        cudaEvent_t e; detail::check_cuda(cudaEventCreate(&e), "start_timer");
        detail::check_cuda(cudaEventRecord(e, stream_), "start_timer");
        if (start_events_.count(operation)) cudaEventDestroy(start_events_[operation]);
        start_events_[operation] = e;
    }
    void stop_timer(const std::string& operation) {
        auto it = start_events_.find(operation);
        if (it == start_events_.end()) throw std::runtime_error("stop_timer: no running timer for '" + operation + "'");
        cudaEvent_t e; detail::check_cuda(cudaEventCreate(&e), "stop_timer");
        detail::check_cuda(cudaEventRecord(e, stream_), "stop_timer");
        intervals_.push_back({operation, it->second, e});
        start_events_.erase(it);
        record_operation(operation);
    }
    void record_operation(const std::string& op_name) { counts_[op_name]++; }
    PerfStats get_stats() const {
        PerfStats s;
        for (const auto& kv : totals()) {
            const std::string& k = kv.first; const double ms = kv.second;
            if (k == "encrypt") s.encrypt_time_ms = ms; else if (k == "decrypt") s.decrypt_time_ms = ms; else if (k == "add") s.add_time_ms = ms;
            else if (k == "multiply" || k == "mul") s.mul_time_ms += ms; else if (k == "relinearize" || k == "relin") s.relin_time_ms += ms;
            else if (k == "mod_switch") s.mod_switch_time_ms = ms; else if (k == "bootstrap") s.bootstrap_time_ms = ms;
        }
        auto cnt = [&](const char* k) { auto it = counts_.find(k); return it == counts_.end() ? (uint64_t)0 : it->second; };
        s.num_encryptions = cnt("encrypt"); s.num_multiplications = cnt("multiply") + cnt("mul"); s.num_bootstraps = cnt("bootstrap");
        return s;
    }
    void print_stats() const {
        for (const auto& kv : totals()) {
            auto it = counts_.find(kv.first);
            std::printf("%-16s %10.3f ms  (%llu)\n", kv.first.c_str(), kv.second, (unsigned long long)(it == counts_.end() ? 0 : it->second));
        }
    }
    void reset() {
        for (auto& kv : start_events_) cudaEventDestroy(kv.second);
        for (auto& iv : intervals_) { cudaEventDestroy(iv.a); cudaEventDestroy(iv.b); }
        start_events_.clear(); intervals_.clear(); counts_.clear();
    }

private:
    struct Interval { std::string op; cudaEvent_t a, b; };
    std::map<std::string, double> totals() const {
        std::map<std::string, double> t;
        for (const auto& iv : intervals_) {
            float ms = 0;
            detail::check_cuda(cudaEventSynchronize(iv.b), "PerformanceMonitor");
            detail::check_cuda(cudaEventElapsedTime(&ms, iv.a, iv.b), "PerformanceMonitor");
            t[iv.op] += ms;
        }
        return t;
    }
    cudaStream_t stream_;
    std::map<std::string, cudaEvent_t> start_events_;
    std::vector<Interval> intervals_;
    std::map<std::string, uint64_t> counts_;
};

}  // namespace fhe
