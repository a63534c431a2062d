// Mirrors the reference's test program (/root/reference/tests/test_fhe.cu) against the compat headers -- same
// parameters and call sequences, but every "expected" value is ASSERTED (the reference only prints them).
// Built and run by tests/test_gpu_compat.py on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "fhe/fhe.cuh"

using namespace fhe;

#define REQUIRE(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)
#define CUDA_CHECK(x) REQUIRE((x) == cudaSuccess)

// the reference declares these in no header and defines them in src/bigint.cu:171-198; tests launch them <<<1,1>>>
__global__ void batch_mod_add_kernel(uint256_t* r, const uint256_t* a, const uint256_t* b, uint256_t m, uint32_t n) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) r[i] = add_mod(a[i], b[i], m);
}
__global__ void batch_mod_sub_kernel(uint256_t* r, const uint256_t* a, const uint256_t* b, uint256_t m, uint32_t n) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) r[i] = sub_mod(a[i], b[i], m);
}

static void test_bigint_arithmetic() {           // tests/test_fhe.cu:24-63
    uint256_t a(12345), b(67890), m(100000), *d_a, *d_b, *d_r, h;
    CUDA_CHECK(cudaMalloc(&d_a, sizeof(uint256_t))); CUDA_CHECK(cudaMalloc(&d_b, sizeof(uint256_t))); CUDA_CHECK(cudaMalloc(&d_r, sizeof(uint256_t)));
    CUDA_CHECK(cudaMemcpy(d_a, &a, sizeof(a), cudaMemcpyHostToDevice)); CUDA_CHECK(cudaMemcpy(d_b, &b, sizeof(b), cudaMemcpyHostToDevice));
    batch_mod_add_kernel<<<1, 1>>>(d_r, d_a, d_b, m, 1);
    CUDA_CHECK(cudaMemcpy(&h, d_r, sizeof(h), cudaMemcpyDeviceToHost));
    REQUIRE(h.limbs[0] == 80235);
    batch_mod_sub_kernel<<<1, 1>>>(d_r, d_a, d_b, m, 1);
    CUDA_CHECK(cudaMemcpy(&h, d_r, sizeof(h), cudaMemcpyDeviceToHost));
    REQUIRE(h.limbs[0] == 44455);
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_r);
    std::printf("bigint arithmetic ok\n");
}

static void test_ntt_transform() {               // tests/test_fhe.cu:65-124
    const uint32_t n = 1024; uint256_t modulus(12289);
    NTTEngine ntt(n, modulus);
    std::vector<uint256_t> h(n), r(n);
    for (uint32_t i = 0; i < n; i++) h[i] = uint256_t(i + 1);
    uint256_t* d; CUDA_CHECK(cudaMalloc(&d, n * sizeof(uint256_t)));
    CUDA_CHECK(cudaMemcpy(d, h.data(), n * sizeof(uint256_t), cudaMemcpyHostToDevice));
    ntt.forward(d); CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(r.data(), d, n * sizeof(uint256_t), cudaMemcpyDeviceToHost));
    bool changed = false; for (uint32_t i = 0; i < n; i++) changed |= r[i].limbs[0] != h[i].limbs[0];
    REQUIRE(changed);
    ntt.inverse(d); CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(r.data(), d, n * sizeof(uint256_t), cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n; i++) REQUIRE(r[i].limbs[0] == h[i].limbs[0] && r[i].limbs[1] == 0);
    cudaFree(d);
    std::printf("ntt round trip ok\n");
}

static void test_polynomial_multiplication() {   // tests/test_fhe.cu:126-167 (+ the check it never makes)
    const uint32_t n = 2048; const uint64_t q = 40961;
    NTTEngine ntt(n, uint256_t(q));
    std::vector<uint256_t> a(n), b(n), r(n);
    std::srand(1);
    for (uint32_t i = 0; i < n; i++) { a[i] = uint256_t(std::rand() % 100); b[i] = uint256_t(std::rand() % 100); }
    uint256_t *d_a, *d_b, *d_r;
    CUDA_CHECK(cudaMalloc(&d_a, n * sizeof(uint256_t))); CUDA_CHECK(cudaMalloc(&d_b, n * sizeof(uint256_t))); CUDA_CHECK(cudaMalloc(&d_r, n * sizeof(uint256_t)));
    CUDA_CHECK(cudaMemcpy(d_a, a.data(), n * sizeof(uint256_t), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(d_b, b.data(), n * sizeof(uint256_t), cudaMemcpyHostToDevice));
    ntt.multiply(d_r, d_a, d_b); CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(r.data(), d_r, n * sizeof(uint256_t), cudaMemcpyDeviceToHost));
    for (uint32_t k = 0; k < n; k++) {           // schoolbook negacyclic product
        int64_t acc = 0;
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t j = (k + n - i) % n;
            const int64_t v = (int64_t)(a[i].limbs[0] * b[j].limbs[0] % q);
            acc += (i <= k) ? v : -v;
        }
        acc %= (int64_t)q; if (acc < 0) acc += q;
        REQUIRE(r[k].limbs[0] == (uint64_t)acc);
    }
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_r);
    std::printf("polynomial multiplication ok\n");
}

static void test_polynomial_mod_switch() {        // include/polynomial.cuh:42,45 (declared only in the reference)
    const uint32_t n = 2048; const uint64_t q = 40961, t = 257;     // q = 1 mod 2n; t coprime to q
    NTTEngine ntt(n, uint256_t(q));
    PolynomialOps ops(n, uint256_t(q), &ntt);
    Polynomial a(n, uint256_t(q)), r(n, uint256_t(t));
    std::vector<uint256_t> h(n);
    for (uint32_t i = 0; i < n; i++) h[i] = uint256_t(((uint64_t)i * 7919u + 13u) % q);
    h[0] = uint256_t(0); h[1] = uint256_t(q - 1); h[2] = uint256_t(q / 2); h[3] = uint256_t(q / 2 + 1);
    CUDA_CHECK(cudaMemcpy(a.coeffs, h.data(), n * sizeof(uint256_t), cudaMemcpyHostToDevice));
    ops.mod_switch(r, a, uint256_t(t));
    CUDA_CHECK(cudaDeviceSynchronize());
    std::vector<uint256_t> o(n);
    CUDA_CHECK(cudaMemcpy(o.data(), r.coeffs, n * sizeof(uint256_t), cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t v = h[i].limbs[0];
        const uint64_t want = ((2 * v * t + q) / (2 * q)) % t;       // round(t/q * v) mod t, exact in integers (q odd: no ties)
        REQUIRE(o[i].limbs[0] == want);
    }
    REQUIRE(r.modulus.limbs[0] == t);
    const double nb = ops.estimate_noise(a);                          // largest centred magnitude is q/2 (coefficients 2 and 3)
    REQUIRE(nb > 14.3 && nb < 14.33);                                 // log2(20480) = 14.32
    std::printf("polynomial mod_switch / estimate_noise ok\n");
}

static void test_rns_context() {
    const uint32_t n = 1024;
    std::vector<uint256_t> primes = generate_rns_primes(40, 3, n);
    RNSContext rns(primes);
    std::vector<uint256_t> x(n), back(n);
    for (uint32_t i = 0; i < n; i++) x[i] = uint256_t(0x123456789abcdefull * (i + 1), i, 0, 0);   // < 2^75 < Q (120 bits)
    uint256_t *d_x, *d_r, *d_s;
    CUDA_CHECK(cudaMalloc(&d_x, n * sizeof(uint256_t))); CUDA_CHECK(cudaMalloc(&d_r, 3 * n * sizeof(uint256_t))); CUDA_CHECK(cudaMalloc(&d_s, 3 * n * sizeof(uint256_t)));
    CUDA_CHECK(cudaMemcpy(d_x, x.data(), n * sizeof(uint256_t), cudaMemcpyHostToDevice));
    rns.to_rns(d_r, d_x, n);
    rns.add_rns(d_s, d_r, d_r, n);       // 2x
    rns.sub_rns(d_s, d_s, d_r, n);       // x
    rns.from_rns(d_x, d_s, n);
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(back.data(), d_x, n * sizeof(uint256_t), cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n; i++) REQUIRE(back[i].limbs[0] == x[i].limbs[0] && back[i].limbs[1] == x[i].limbs[1] && back[i].limbs[2] == 0);
    cudaFree(d_x); cudaFree(d_r); cudaFree(d_s);
    std::printf("rns context ok\n");
}

static void test_fhe_operations() {              // tests/test_fhe.cu:169-273
    SecurityParams params;
    params.lambda = 128; params.poly_degree = 4096; params.log_q = 120; params.sigma = 3.2f; params.hamming_weight = 64;
    FHEContext ctx(params);
    REQUIRE(ctx.params().L == 2);
    PublicKey pk; SecretKey sk; RelinKeys rlk;
    ctx.keygen(pk, sk);
    ctx.relinkey_gen(rlk, sk, 16);
    std::vector<uint64_t> v1 = {5, 10, 15, 20}, v2 = {3, 6, 9, 12};
    Plaintext pt1, pt2; ctx.encode(pt1, v1); ctx.encode(pt2, v2);
    Ciphertext ct1, ct2; ctx.encrypt(ct1, pt1, pk); ctx.encrypt(ct2, pt2, pk);
    Ciphertext ct_add; ctx.add(ct_add, ct1, ct2);
    Ciphertext ct_mul; ctx.multiply(ct_mul, ct1, ct2, rlk);
    Plaintext pa, pm, p1; ctx.decrypt(pa, ct_add, sk); ctx.decrypt(pm, ct_mul, sk); ctx.decrypt(p1, ct1, sk);
    std::vector<uint64_t> ra, rm, r1; ctx.decode(ra, pa); ctx.decode(rm, pm); ctx.decode(r1, p1);
    const uint64_t exp_add[4] = {8, 16, 24, 32};                    // tests/test_fhe.cu:264
    const uint64_t exp_mul[8] = {15, 60, 150, 300, 375, 360, 240, 0};  // coefficient encoding -> negacyclic product
    for (int i = 0; i < 4; i++) { REQUIRE(ra[i] == exp_add[i]); REQUIRE(r1[i] == v1[i]); }
    for (int i = 4; i < 4096; i++) { REQUIRE(ra[i] == 0); REQUIRE(r1[i] == 0); }
    for (int i = 0; i < 8; i++) REQUIRE(rm[i] == exp_mul[i]);
    for (int i = 8; i < 4096; i++) REQUIRE(rm[i] == 0);
    // chained: (ct1 * ct2) * ct1, result written over an operand
    ctx.multiply(ct_mul, ct_mul, ct1, rlk);
    Plaintext pc; ctx.decrypt(pc, ct_mul, sk); std::vector<uint64_t> rc; ctx.decode(rc, pc);
    const uint64_t exp_chain[4] = {75, 450, 1575, 4200};            // (15+60x+150x^2+300x^3+..)(5+10x+15x^2+20x^3) low coefficients
    for (int i = 0; i < 4; i++) REQUIRE(rc[i] == exp_chain[i]);
    // SIMD slot encoding: the reference's printed expectations (tests/test_fhe.cu:270; examples/homomorphic_operations.cu:148,242)
    {
        BatchEncoder enc(ctx);
        REQUIRE(enc.slot_count() == 2048);
        Plaintext b1, b2, two; enc.encode(b1, v1); enc.encode(b2, v2); enc.encode(two, std::vector<uint64_t>(4096, 2));
        Ciphertext e1, e2, prod, sum, diff, scaled; ctx.encrypt(e1, b1, pk); ctx.encrypt(e2, b2, pk);
        ctx.multiply(prod, e1, e2, rlk); ctx.add_plain(sum, e1, b2); ctx.sub(diff, e1, e2); ctx.multiply_plain(scaled, e1, two);
        Plaintext pp, ps, pd, pq; ctx.decrypt(pp, prod, sk); ctx.decrypt(ps, sum, sk); ctx.decrypt(pd, diff, sk); ctx.decrypt(pq, scaled, sk);
        std::vector<uint64_t> rp, rs, rd, rq; enc.decode(rp, pp); enc.decode(rs, ps); enc.decode(rd, pd); enc.decode(rq, pq);
        const uint64_t exp_slot[4] = {15, 60, 135, 240};
        for (int i = 0; i < 4; i++) { REQUIRE(rp[i] == exp_slot[i]); REQUIRE(rs[i] == v1[i] + v2[i]); REQUIRE(rd[i] == v1[i] - v2[i]); REQUIRE(rq[i] == 2 * v1[i]); }
        for (int i = 4; i < 4096; i++) REQUIRE(rp[i] == 0);
        FHEContext::release(e1); FHEContext::release(e2); FHEContext::release(prod); FHEContext::release(sum); FHEContext::release(diff); FHEContext::release(scaled);
        FHEContext::release(b1); FHEContext::release(b2); FHEContext::release(two); FHEContext::release(pp); FHEContext::release(ps); FHEContext::release(pd); FHEContext::release(pq);
    }
    // noise budget (include/fhe.cuh:142): log2 Q = 120 bits; a fresh ciphertext keeps most of it, a product less, a chain less still
    {
        const float fresh = ctx.estimate_noise_budget(ct1, sk);
        Ciphertext once; ctx.multiply(once, ct1, ct2, rlk);
        const float after1 = ctx.estimate_noise_budget(once, sk), after2 = ctx.estimate_noise_budget(ct_mul, sk);
        REQUIRE(fresh > 70.0f && fresh < 119.0f);
        REQUIRE(after1 > 20.0f && after1 < fresh - 10.0f);
        REQUIRE(after2 > 0.0f && after2 < after1);
        FHEContext::release(once);
    }
    // multiply and relinearize as separate calls (src/fhe.cu:198-235): a sum of two 3-component products relinearised once,
    // and the square of a ciphertext (same operand twice)
    {
        Ciphertext p12, p11, sum3, sq;
        ctx.multiply(p12, ct1, ct2); ctx.multiply(p11, ct1, ct1);
        REQUIRE(p12.components.size() == 3 && p11.components.size() == 3);
        ctx.add(sum3, p12, p11);
        REQUIRE(sum3.components.size() == 3);
        ctx.relinearize(sum3, rlk);
        REQUIRE(sum3.components.size() == 2);
        ctx.relinearize(sum3, rlk);                                  // two components: a no-op, as in the reference
        ctx.multiply(sq, ct1, ct1, rlk);
        Plaintext ps3, psq; ctx.decrypt(ps3, sum3, sk); ctx.decrypt(psq, sq, sk);
        std::vector<uint64_t> r3, rq; ctx.decode(r3, ps3); ctx.decode(rq, psq);
        const uint64_t exp_sq[8] = {25, 100, 250, 500, 625, 600, 400, 0};       // (5 + 10x + 15x^2 + 20x^3)^2
        for (int i = 0; i < 8; i++) { REQUIRE(rq[i] == exp_sq[i]); REQUIRE(r3[i] == exp_mul[i] + exp_sq[i]); }
        for (int i = 8; i < 4096; i++) { REQUIRE(rq[i] == 0); REQUIRE(r3[i] == 0); }
        FHEContext::release(p12); FHEContext::release(p11); FHEContext::release(sum3); FHEContext::release(sq);
        FHEContext::release(ps3); FHEContext::release(psq);
    }
    // PerformanceMonitor (include/fhe.cuh:169-198, declared only in the reference): CUDA-event timing on the context's stream
    {
        PerformanceMonitor mon(ctx.stream());
        Ciphertext tmp;
        for (int i = 0; i < 3; i++) { mon.start_timer("multiply"); ctx.multiply(tmp, ct1, ct2, rlk); mon.stop_timer("multiply"); }
        mon.start_timer("encrypt"); Ciphertext e3; ctx.encrypt(e3, pt1, pk); mon.stop_timer("encrypt");
        const PerfStats st = mon.get_stats();
        REQUIRE(st.num_multiplications == 3 && st.num_encryptions == 1);
        REQUIRE(st.mul_time_ms > 0.0 && st.mul_time_ms < 1000.0 && st.encrypt_time_ms > 0.0);
        bool threw = false;
        try { mon.stop_timer("never started"); } catch (const std::runtime_error&) { threw = true; }
        REQUIRE(threw);
        mon.reset();
        REQUIRE(mon.get_stats().num_multiplications == 0);
        FHEContext::release(tmp); FHEContext::release(e3);
    }
    // rotations (include/fhe.cuh:113-116, declared only in the reference): the slot at NTT position k holds the evaluation at
    // psi^(2 bitrev(k) + 1); x -> x^g sends the value of position k' to position k with exponent(k') = exponent(k) * g mod 2N
    {
        BatchEncoder enc(ctx);
        GaloisKeys gk; ctx.galoiskey_gen(gk, sk);
        REQUIRE(gk.gal_keys.size() == 12);                          // 3^(2^k), k = 0..10, and 2N - 1
        const uint32_t n = 4096, lg = 12;
        std::vector<uint64_t> vals(n); for (uint32_t i = 0; i < n; i++) vals[i] = (i * 7 + 3) % 65537;
        Plaintext bp; enc.encode(bp, vals);
        Ciphertext e, r1, r5, r6, cc, c2; ctx.encrypt(e, bp, pk);
        ctx.rotate_rows(r1, e, 1, gk); ctx.rotate_rows(r5, e, 5, gk); ctx.rotate_rows(r6, r5, 1, gk);
        ctx.rotate_columns(cc, e, gk); ctx.rotate_columns(c2, cc, gk);
        auto slots = [&](Ciphertext& c) { Plaintext p; ctx.decrypt(p, c, sk); std::vector<uint64_t> v; enc.decode(v, p); FHEContext::release(p); return v; };
        auto brev = [&](uint32_t k) { uint32_t r = 0; for (uint32_t b = 0; b < lg; b++) r |= ((k >> b) & 1u) << (lg - 1 - b); return r; };
        std::vector<uint32_t> pos_of_exp(2 * n, 0);
        for (uint32_t k = 0; k < n; k++) pos_of_exp[2 * brev(k) + 1] = k;
        auto expect = [&](uint64_t g) { std::vector<uint64_t> v(n); for (uint32_t k = 0; k < n; k++) v[k] = vals[pos_of_exp[(uint32_t)(((2ull * brev(k) + 1) * g) % (2 * n))]]; return v; };
        uint64_t g5 = 1; for (int i = 0; i < 5; i++) g5 = g5 * 3 % (2 * n);
        REQUIRE(slots(r1) == expect(3));
        REQUIRE(slots(r5) == expect(g5));
        REQUIRE(slots(r6) == expect(g5 * 3 % (2 * n)));             // rotations compose
        REQUIRE(slots(cc) == expect(2 * n - 1));
        REQUIRE(slots(c2) == vals);                                 // the column swap is an involution
        FHEContext::release(e); FHEContext::release(r1); FHEContext::release(r5); FHEContext::release(r6); FHEContext::release(cc); FHEContext::release(c2);
        FHEContext::release(bp); FHEContext::release(gk);
    }
    FHEContext::release(ct1); FHEContext::release(ct2); FHEContext::release(ct_add); FHEContext::release(ct_mul);
    FHEContext::release(pt1); FHEContext::release(pt2); FHEContext::release(pa); FHEContext::release(pm); FHEContext::release(p1); FHEContext::release(pc);
    FHEContext::release(rlk); FHEContext::release(pk); FHEContext::release(sk);
    std::printf("fhe operations ok\n");
}

static void test_errors() {
    bool threw = false;
    try { NTTEngine bad(4096, uint256_t(12289)); } catch (const std::runtime_error&) { threw = true; }   // 12289 != 1 mod 8192
    REQUIRE(threw);
    std::printf("error behaviour ok\n");
}

int main() {
    test_bigint_arithmetic();
    test_ntt_transform();
    test_polynomial_multiplication();
    test_polynomial_mod_switch();
    test_rns_context();
    test_fhe_operations();
    test_errors();
    std::printf("ALL COMPAT TESTS PASSED\n");
    return 0;
}
