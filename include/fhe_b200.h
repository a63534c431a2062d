/*
 * fhe_b200.h -- C ABI of the B200-native RNS-NTT polynomial-arithmetic engine (libfhe_b200.so).
 *
 * This is the drop-in boundary for the hot path under fhe::FHEContext multiply / relinearize / encrypt of
 * codebasecomprehension987/gpu-homomorphic-encryption.  The reference has no FFI layer (it is a C++ static
 * library, CMakeLists.txt:29-30); each entry point below names the reference interface it replaces.  The
 * compat C++ headers in include/fhe/ re-implement the reference's fhe:: classes purely on top of this file.
 *
 * Conventions
 *   - every function returns 0 on success, a negative FHE_B200_E* code on failure; fhe_b200_last_error() gives
 *     the message for the calling thread.  Nothing throws, no C++ type crosses the boundary.
 *   - pointers named d_* are DEVICE pointers owned by the caller, h_* are host pointers.  Device buffers of polynomials must be
 *     16-byte aligned (the kernels use 16-byte accesses and bulk copies; anything from cudaMalloc, or offset by whole limbs, is).
 *   - `stream` is a cudaStream_t passed as void*; all device work is asynchronous and ordered on that stream.  BFV calls on
 *     two or more ciphertexts (and a single multiply) fork part of their work onto streams owned by the context and join it
 *     back before the call's work completes on `stream`, so callers need no extra synchronisation (FHE_B200_HMULT_STREAMS=1
 *     keeps everything on `stream`).  A BFV context owns one workspace: its calls must be stream-ordered, not concurrent.
 *   - objects run on the device they were created for; a call switches to it and restores the caller's current device.
 *   - polynomial layout is limb-major uint64_t [batch][limb_count][N]; limb l of every polynomial is reduced
 *     modulo plan.moduli[limb_begin + l]; values are canonical residues in [0, q) at every call boundary.
 *   - NTT-domain order: forward leaves X[k] = sum_j a_j psi^(j(2k+1)) at position bitrev(k); inverse consumes
 *     that order.  Pointwise operations are order-agnostic.  fhe_b200_bitrev_permute converts to natural order.
 *   - a plan (and a BFV context) must not be used from two host threads at once (the reference documents the
 *     same rule, docs/API_REFERENCE.md:600-606).
 *   - there is no CPU fallback: every entry point needs a CUDA device of compute capability 10.0.
 */
#ifndef FHE_B200_H
#define FHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FHE_B200_OK 0
#define FHE_B200_EINVAL (-1)   /* bad argument (N not a power of two in [512, 131072], modulus not NTT friendly, ...) */
#define FHE_B200_ECUDA (-2)    /* a CUDA runtime call failed */
#define FHE_B200_ENOMEM (-3)
#define FHE_B200_ESTATE (-4)   /* object used in the wrong state */

typedef struct fhe_b200_plan fhe_b200_plan;
typedef struct fhe_b200_lincomb fhe_b200_lincomb;
typedef struct fhe_b200_bfv fhe_b200_bfv;

const char* fhe_b200_last_error(void);
int fhe_b200_version(void);
/* number of kernels this library has launched in this process (all threads) */
uint64_t fhe_b200_launch_count(void);
/* per-kernel timing for the roofline report: while enabled, every NTT kernel launch is bracketed by CUDA events
 * on its own stream.  kind: 0 = tile pass (forward), 1 = tile pass (inverse), 2 = row pass (forward), 3 = row pass
 * (inverse), 4 = RNS linear combination, 5 = BFV tensor product, 6 = key-switch inner product.  profile_read synchronises the recorded events and returns launches, total milliseconds and the
 * limb-transforms those launches covered. */
int fhe_b200_profile_enable(int on);
int fhe_b200_profile_read(int kind, uint64_t* launches, double* total_ms, uint64_t* limb_transforms);

/* Integer-pipe peaks of `device`, measured now (csrc/peaks.cu): thread-level operations per second of IMAD (mad.lo.u32) and of
 * IMAD.WIDE (mul.wide.u32) over the whole chip, best of `reps` launches each.  The butterfly of the transforms is bound by this
 * pipe, and bench.py uses the two figures as its roofline denominators.  Synchronous; nothing in the reference corresponds. */
int fhe_b200_measure_int_peaks(int device, double* imad_lo_ops, double* imad_wide_ops, int reps);
/* The transforms' own lazy butterfly (same multiply, bounds and range reductions) in a loop that never leaves the registers: lazy
 * butterflies per second over the whole chip, best of `reps` launches -- the practical ceiling of the instruction mix, which bench.py
 * prints beside the IMAD roofline.  q: a modulus in (2^60 - 2^32, 2^60); w < q.  Synchronous. */
int fhe_b200_measure_butterfly_loop(int device, uint64_t q, uint64_t w, double* butterflies_per_s, int reps);

/* ---- plan: N, the RNS moduli and their device-resident twiddle tables ------------------------------------
 * replaces NTTEngine::NTTEngine / precompute_twiddle_factors / find_primitive_root / mod_inverse
 * (src/ntt.cu:7-22,77-119) and RNS_NTTEngine::RNS_NTTEngine (src/ntt.cu:122-141).
 * Every modulus must be an odd prime < 2^61 with q = 1 (mod 2N). */
int fhe_b200_plan_create(uint32_t n, const uint64_t* h_moduli, uint32_t n_limbs, int device, fhe_b200_plan** out);
int fhe_b200_plan_destroy(fhe_b200_plan* plan);
uint32_t fhe_b200_plan_n(const fhe_b200_plan* plan);
uint32_t fhe_b200_plan_limbs(const fhe_b200_plan* plan);
int fhe_b200_plan_moduli(const fhe_b200_plan* plan, uint64_t* h_out);
/* copies limb `limb`'s tables to the host: w[N] and shoup[N] for each direction (for parity tests) */
int fhe_b200_plan_tables(const fhe_b200_plan* plan, uint32_t limb, uint64_t* h_fwd, uint64_t* h_fwd_shoup,
                         uint64_t* h_inv, uint64_t* h_inv_shoup);

/* ---- transforms -------------------------------------------------------------------------------------------
 * replaces NTTEngine::forward / inverse / forward_batch / inverse_batch (src/ntt.cu:30-47, include/ntt.cuh:87-88),
 * RNS_NTTEngine::forward_rns / inverse_rns (src/ntt.cu:158-171) and the kernels in kernels/ntt_kernels.cu.
 * d_out may equal d_in (in place).  Layout [batch][limb_count][N]. */
int fhe_b200_ntt_forward(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t batch,
                         uint32_t limb_begin, uint32_t limb_count, void* stream);
int fhe_b200_ntt_inverse(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t batch,
                         uint32_t limb_begin, uint32_t limb_count, void* stream);
/* out = a * b in Z_q[x]/(x^N+1) per limb, coefficient form in and out.
 * replaces NTTEngine::multiply (src/ntt.cu:49-75), RNS_NTTEngine::multiply_rns (include/ntt.cuh:123-126),
 * PolynomialOps::mul_ntt / mul_negacyclic (src/polynomial.cu:54-58, include/polynomial.cuh:36-39). */
int fhe_b200_negacyclic_mul(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_a, const uint64_t* d_b,
                            uint32_t batch, uint32_t limb_begin, uint32_t limb_count, void* stream);
/* out[bitrev(k)] = in[k] per limb (out != in) -- the job of bit_reverse_kernel (kernels/ntt_kernels.cu:140-161),
 * only needed by callers that want natural-order NTT values. */
int fhe_b200_bitrev_permute(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t n_polys,
                            void* stream);
/* host-buffer variant: copies h_data (pinned or pageable) to the device in pipelined chunks, transforms, copies
 * back.  direction: 0 forward, 1 inverse, 2 forward then inverse.  Synchronous. */
int fhe_b200_ntt_host(fhe_b200_plan* plan, uint64_t* h_data, uint32_t batch, uint32_t limb_begin,
                      uint32_t limb_count, int direction);

/* ---- element-wise ------------------------------------------------------------------------------------------
 * replaces poly_add / poly_sub / poly_mul_scalar kernels (src/polynomial.cu:70-111), ntt_pointwise_mul_kernel
 * (kernels/ntt_kernels.cu:124-137), rns_add / rns_sub / rns_mul kernels (src/rns.cu:143-180, include/rns.cuh:96-103). */
int fhe_b200_poly_add(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_a, const uint64_t* d_b,
                      uint32_t batch, uint32_t limb_begin, uint32_t limb_count, void* stream);
int fhe_b200_poly_sub(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_a, const uint64_t* d_b,
                      uint32_t batch, uint32_t limb_begin, uint32_t limb_count, void* stream);
int fhe_b200_poly_mul(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_a, const uint64_t* d_b,
                      uint32_t batch, uint32_t limb_begin, uint32_t limb_count, void* stream);
/* out = acc + a*b */
int fhe_b200_poly_mac(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_acc, const uint64_t* d_a,
                      const uint64_t* d_b, uint32_t batch, uint32_t limb_begin, uint32_t limb_count, void* stream);
/* out = a * scalar; h_scalars[limb_count] holds the scalar's residue for each limb */
int fhe_b200_poly_mul_scalar(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_a, const uint64_t* h_scalars,
                             uint32_t batch, uint32_t limb_begin, uint32_t limb_count, void* stream);
int fhe_b200_poly_add_scalar(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_a, const uint64_t* h_scalars,
                             uint32_t batch, uint32_t limb_begin, uint32_t limb_count, void* stream);
int fhe_b200_poly_negate(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_a, uint32_t batch,
                         uint32_t limb_begin, uint32_t limb_count, void* stream);

/* ---- fhe::uint256_t edge conversion (include/bigint.cuh:9-24: four little-endian 64-bit words) ---------------
 * the reference API passes 32-byte coefficients; the engine works on 8-byte residues.  unpack keeps limbs[0]
 * (all hot-path values are < 2^61); pack zero-extends. */
int fhe_b200_unpack_u256(uint64_t* d_out, const void* d_u256, size_t count, void* stream);
int fhe_b200_pack_u256(void* d_u256, const uint64_t* d_in, size_t count, void* stream);
/* RNSContext::to_rns (src/rns.cu:57-63): residues of `count` 256-bit values for limbs [limb_begin, +limb_count),
 * written limb-major [limb_count][count]. */
int fhe_b200_to_rns_u256(fhe_b200_plan* plan, uint64_t* d_out, const void* d_u256, size_t count,
                         uint32_t limb_begin, uint32_t limb_count, void* stream);

/* RNSContext::from_rns / RNS_NTTEngine::from_rns (src/rns.cu:65-72, from_rns_crt_kernel :117-141 -- writes 0 in the
 * reference): CRT reconstruction of `count` values from limb-major residues [limb_count][count] into 256-bit words,
 * value in [0, Q).  Needs Q = prod q_i < 2^256 (at most four 60-bit limbs). */
int fhe_b200_from_rns_u256(fhe_b200_plan* plan, void* d_u256, const uint64_t* d_in, size_t count, uint32_t limb_begin,
                           uint32_t limb_count, void* stream);

/* ---- RNS linear combination: exact base conversion and t/Q scale-and-round ----------------------------------
 * replaces fast_base_conversion_kernel / RNSContext::base_extend (include/rns.cuh:47-49,116-125),
 * rns_mod_switch_kernel / mod_switch_rns (include/rns.cuh:44-45,128-136), poly_mod_switch_kernel
 * (include/polynomial.cuh:96-102) and from_rns_crt_kernel's role in decryption (src/rns.cu:117-141).
 *   out_k = ( sum_i z_i M[i][k] + (I mod m_k) c_k + extra_k lam_k ) mod m_k ,
 *   z_i = x_i pre_i mod s_i (conversion) or x_i (scaling),  I = round( sum_i z_i theta_i ) in 128-bit fixed point.
 * in: [batch][S][N]   extra: [batch][T][N] or NULL   out: [batch][T][N]     (n_coeffs = N of the caller's ring) */
int fhe_b200_lincomb_create_conv(const uint64_t* h_src, uint32_t S, const uint64_t* h_dst, uint32_t T, int device,
                                 fhe_b200_lincomb** out);
int fhe_b200_lincomb_create_scale(const uint64_t* h_q, uint32_t L, const uint64_t* h_p, uint32_t R, uint64_t t,
                                  const uint64_t* h_targets, uint32_t T, int with_extra, int device,
                                  fhe_b200_lincomb** out);   /* with_extra: every target must be one of h_p (any subset, any order) */
int fhe_b200_lincomb_destroy(fhe_b200_lincomb* lc);
int fhe_b200_lincomb_apply(fhe_b200_lincomb* lc, uint64_t* d_out, const uint64_t* d_in, const uint64_t* d_extra,
                           uint32_t n_coeffs, uint32_t batch, void* stream);
/* copies the precomputed constants to the host (for parity tests): pre[S], th_hi[S], th_lo[S], M[S*T], c[T], lam[T] */
int fhe_b200_lincomb_constants(const fhe_b200_lincomb* lc, uint64_t* pre, uint64_t* th_hi, uint64_t* th_lo,
                               uint64_t* M, uint64_t* c, uint64_t* lam);
/* drop the last limb with rounding: out_i = round(x / q_last) mod q_i, in [batch][limbs][N] -> out [batch][limbs-1][N] */
int fhe_b200_modswitch_drop_last(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t batch,
                                 uint32_t limb_begin, uint32_t limb_count, void* stream);

/* ---- BFV on RNS: the path under fhe::FHEContext ------------------------------------------------------------
 * moduli = Q (L primes) followed by the auxiliary basis (R primes); the special modulus of hybrid key switching
 * is the first K auxiliary primes; dnum digits of alpha = L/dnum limbs each.
 * Device data formats (uint64_t):
 *   secret key   [L+R][N]            NTT form          public key  [2][L][N]   NTT form  (pk0 = e - a s, pk1 = a)
 *   relin key    [dnum][2][L+K][N]   NTT form          ciphertext  [2][L][N]   coefficient form
 *   plaintext    [N]                 coefficients mod t
 * Randomness is a counter-based generator keyed by the caller's 64-bit seed (specification in DESIGN.md), so
 * results are reproducible and identical to the CPU oracle's. */
int fhe_b200_bfv_create(uint32_t n, uint32_t L, uint32_t R, uint32_t K, uint32_t dnum, uint64_t t,
                        const uint64_t* h_moduli, float sigma, uint32_t hamming_weight, int device,
                        fhe_b200_bfv** out);
int fhe_b200_bfv_destroy(fhe_b200_bfv* ctx);
/* Randomness.  By default every sampler draws from a reproducible counter hash (splitmix64) keyed by the call's 64-bit seed: right
 * for tests and benchmarks, NOT for keys that matter.  fhe_b200_bfv_set_rng_key gives the context a 256-bit key (take it from the
 * OS); from then on every random word is ChaCha20(key; nonce = call seed and stream; counter = word index / 8) and the 64-bit seeds
 * only separate calls.  NULL switches back.  (The reference's samplers are placeholders seeded by rand(), src/fhe.cu:238-257.) */
int fhe_b200_bfv_set_rng_key(fhe_b200_bfv* ctx, const uint8_t* h_key32);
/* FHEContext::keygen (src/fhe.cu:54-74) */
int fhe_b200_bfv_keygen(fhe_b200_bfv* ctx, uint64_t seed_sk, uint64_t seed_pk, uint64_t* d_sk, uint64_t* d_pk,
                        void* stream);
/* FHEContext::relinkey_gen (src/fhe.cu:76-111) */
int fhe_b200_bfv_relinkeygen(fhe_b200_bfv* ctx, uint64_t seed, const uint64_t* d_sk, uint64_t* d_rlk, void* stream);
/* FHEContext::encrypt (src/fhe.cu:138-169); batch plaintexts [batch][N] -> ciphertexts [batch][2][L][N]; ciphertext b is batch item b
 * of the call: its generator seed is a hash of (seed, b), see DESIGN.md "Randomness" */
int fhe_b200_bfv_encrypt(fhe_b200_bfv* ctx, uint64_t seed, const uint64_t* d_pt, const uint64_t* d_pk,
                         uint64_t* d_ct, uint32_t batch, void* stream);
/* FHEContext::decrypt (src/fhe.cu:171-185) */
int fhe_b200_bfv_decrypt(fhe_b200_bfv* ctx, const uint64_t* d_ct, const uint64_t* d_sk, uint64_t* d_pt,
                         uint32_t batch, void* stream);
/* FHEContext::add (src/fhe.cu:187-197) */
int fhe_b200_bfv_add(fhe_b200_bfv* ctx, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out, uint32_t batch,
                     void* stream);
/* FHEContext::sub, add_plain, sub_plain, multiply_plain (declared include/fhe.cuh:98-104, undefined in the reference).
 * Plaintexts are [batch][N] coefficients mod t.  add/sub_plain: c0 +- Delta*m.  multiply_plain: (c0*m, c1*m) in R_Q with m
 * lifted from [0,t). */
int fhe_b200_bfv_sub(fhe_b200_bfv* ctx, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out, uint32_t batch,
                     void* stream);
int fhe_b200_bfv_add_plain(fhe_b200_bfv* ctx, const uint64_t* d_ct, const uint64_t* d_pt, uint64_t* d_out, uint32_t batch,
                           int subtract, void* stream);
int fhe_b200_bfv_multiply_plain(fhe_b200_bfv* ctx, const uint64_t* d_ct, const uint64_t* d_pt, uint64_t* d_out,
                                uint32_t batch, void* stream);
/* FHEContext::multiply + relinearize (src/fhe.cu:199-235).  d_scaled (optional, [batch][3][L][N]) receives the
 * scaled tensor before relinearisation.  For batch >= 2 the work is forked from `stream` onto two internal streams (one half of the
 * batch each) and joined back before the call's work is complete on `stream`; FHE_B200_HMULT_STREAMS=1 disables the split. */
int fhe_b200_bfv_multiply_relin(fhe_b200_bfv* ctx, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_rlk,
                                uint64_t* d_out, uint64_t* d_scaled, uint32_t batch, void* stream);
/* FHEContext::estimate_noise_budget (include/fhe.cuh:142, declared only): invariant noise budget in bits,
 * log2(Q) - log2(||[t (c0 + c1 s)]_Q||_inf) - 1, one value per ciphertext into the HOST array h_bits[batch].  The device
 * computes t (c0 + c1 s) per limb; the host centres every coefficient through mixed-radix digits.  Synchronises the stream;
 * a diagnostic (O(N L^2) host work), not a hot-path call. */
int fhe_b200_bfv_noise_budget(fhe_b200_bfv* ctx, const uint64_t* d_ct, const uint64_t* d_sk, uint32_t batch, double* h_bits,
                              void* stream);
/* The two halves on their own, as the reference's API has them (multiply builds three components, src/fhe.cu:198-219;
 * relinearize reduces them to two, :226-235 -- a stub there): d_out3 and d_ct3 are [batch][3][L][N] coefficient form.  Sums of
 * 3-component ciphertexts (fhe_b200_poly_add over 3*batch polynomials) can be relinearised once.  d_a == d_b (same pointer) takes
 * a squaring path in both multiply entry points (two polynomials extended and transformed instead of four; same words).
 * relinearize: d_out [batch][2][L][N] must not alias d_ct3. */
int fhe_b200_bfv_multiply(fhe_b200_bfv* ctx, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out3, uint32_t batch,
                          void* stream);
int fhe_b200_bfv_relinearize(fhe_b200_bfv* ctx, const uint64_t* d_ct3, const uint64_t* d_rlk, uint64_t* d_out, uint32_t batch,
                             void* stream);
/* Building blocks of multiply_relin on a limb range, for limb-sharded execution (one rank owns limbs
 * [limb_begin, limb_begin+limb_count) of every polynomial; NTT form in, NTT form out).
 *   tensor  : ext [4][batch][limb_count][N] = (a0,a1,b0,b1)  ->  out [3][batch][limb_count][N] = (a0 b0, a0 b1 + a1 b0, a1 b1)
 *   ks_inner: acc[c][b][i] = sum_d dig[d][b][i] * key[d][c][i],  dig [dnum][batch][limb_count][N], key [dnum][2][limb_count][N] */
int fhe_b200_bfv_tensor(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_ext, uint32_t batch, uint32_t limb_begin,
                        uint32_t limb_count, void* stream);
int fhe_b200_bfv_ks_inner(fhe_b200_plan* plan, uint64_t* d_acc, const uint64_t* d_dig, const uint64_t* d_key, uint32_t dnum,
                          uint32_t batch, uint32_t limb_begin, uint32_t limb_count, void* stream);
/* Galois keys and automorphisms: FHEContext::galoiskey_gen / rotate_rows / rotate_columns (include/fhe.cuh:59-61,86,113-116;
 * declared only in the reference).  galois_elt g is odd and < 2N; the key has the relinearisation-key layout
 * [dnum][2][L+K][N] (NTT form) and switches s(x^g) back to s.  apply_galois: ct(x) -> ct(x^g) re-encrypted under s,
 * ciphertexts [batch][2][L][N] coefficient form.  rotate_rows(steps) = apply_galois(3^steps mod 2N), rotate_columns =
 * apply_galois(2N - 1) (compat layer). */
int fhe_b200_bfv_galoiskeygen(fhe_b200_bfv* ctx, uint64_t seed, uint32_t galois_elt, const uint64_t* d_sk, uint64_t* d_gk,
                              void* stream);
int fhe_b200_bfv_apply_galois(fhe_b200_bfv* ctx, const uint64_t* d_ct, uint32_t galois_elt, const uint64_t* d_gk,
                              uint64_t* d_out, uint32_t batch, void* stream);
/* FHEContext::mod_switch_to_next (include/fhe.cuh:109; declared only): both components lose the last limb of Q with rounding,
 * [batch][2][L][N] -> [batch][2][L-1][N]; the result decrypts under a context built on the first L-1 limbs. */
int fhe_b200_bfv_mod_switch_to_next(fhe_b200_bfv* ctx, const uint64_t* d_ct, uint64_t* d_out, uint32_t batch, void* stream);
/* FHEContext::mod_switch_to_level (include/fhe.cuh:110, declared only): `drop` limbs are removed one at a time (each step as
 * mod_switch_to_next); d_out is [batch][2][L-drop][N] and belongs to the context on the first L-drop limbs.  drop = 0 copies. */
int fhe_b200_bfv_mod_switch_to_level(fhe_b200_bfv* ctx, const uint64_t* d_ct, uint64_t* d_out, uint32_t batch, uint32_t drop,
                                     void* stream);
/* host-buffer variant of multiply_relin (copies in, computes, copies out; synchronous) */
int fhe_b200_bfv_multiply_relin_host(fhe_b200_bfv* ctx, const uint64_t* h_a, const uint64_t* h_b,
                                     const uint64_t* d_rlk, uint64_t* h_out, uint32_t batch);
/* ---- limb-sharded multiply + relinearize over the GPUs of one NVLink domain (BASELINE.json config 4) -----------------------
 * The reference only sketches multi-GPU execution ("limbs distributed over GPUs", docs/ARCHITECTURE.md:501-512, README.md:320).
 * One shard object per GPU (rank) on top of that GPU's own BFV context (same parameters on every rank, world a power of two
 * <= 16, N/world >= 256).  NTT-domain work runs limb-sharded (rank g owns a block of the L+R / L+K limbs), base conversions run
 * coefficient-sharded (rank g owns coefficients [g N/world, (g+1) N/world)); the four transpositions per multiply are stores of
 * the producing kernels straight into the peer's buffer over NVLink, ordered by epoch flags (csrc/shard.cu) -- no NCCL call and
 * no host synchronisation on the data path.  Set-up: create on every rank, export the 128-byte handle, carry all handles to
 * every rank in rank order (torch.distributed.all_gather_object, MPI, a pipe ...), connect.  Ranks may be processes (CUDA IPC)
 * or several objects in one process (peer access; also several ranks on ONE device, which is how the single-GPU tests run it --
 * stage by stage, see fhe_b200_bfv_multiply_relin_sharded_stage).
 * Data: sharded ciphertext [batch][2][L][N/world] = the rank's coefficient block of every limb, coefficient form;
 *       key slice [dnum][2][cW][N] = the rank's key limbs (fhe_b200_shard_slice_key).
 * Every rank must issue the same sequence of sharded calls.  A rank whose peer never delivers gives up after
 * FHE_B200_SHARD_TIMEOUT_MS (default 10 000) per wait and fhe_b200_shard_check reports it. */
typedef struct fhe_b200_shard fhe_b200_shard;
#define FHE_B200_SHARD_HANDLE_BYTES 128
int fhe_b200_shard_create(fhe_b200_bfv* ctx, int rank, int world, uint32_t max_batch, fhe_b200_shard** out);
int fhe_b200_shard_destroy(fhe_b200_shard* shard);
int fhe_b200_shard_handle(const fhe_b200_shard* shard, void* h_handle);
int fhe_b200_shard_connect(fhe_b200_shard* shard, const void* h_handles /* [world][128], rank order */);
/* the block partition used for limbs: the first total % world ranks own one item more (host only, no GPU needed) */
int fhe_b200_shard_partition(uint32_t total, uint32_t rank, uint32_t world, uint32_t* begin, uint32_t* count);
/* this rank's coefficient block, key limbs (of L+K) and extended-basis limbs (of L+R); bytes it stores into other ranks per
 * multiply of one ciphertext pair (any output may be NULL) */
int fhe_b200_shard_info(const fhe_b200_shard* shard, uint32_t* coeff_begin, uint32_t* coeff_count, uint32_t* key_limb_begin,
                        uint32_t* key_limb_count, uint32_t* ext_limb_begin, uint32_t* ext_limb_count, uint64_t* nvlink_bytes_per_op);
int fhe_b200_shard_slice_key(const fhe_b200_shard* shard, const uint64_t* d_key, uint64_t* d_slice, void* stream);
/* FHEContext::multiply + relinearize (src/fhe.cu:199-235) on sharded ciphertexts; same words as fhe_b200_bfv_multiply_relin */
int fhe_b200_bfv_multiply_relin_sharded(fhe_b200_shard* shard, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_key_slice,
                                        uint64_t* d_out, uint32_t batch, void* stream);
/* One stage (0..4) of the same multiply: stage k starts with the wait for the peers' phase k-1 and ends with this rank's signal of
 * phase k.  Only for a host that drives several ranks on ONE device (the single-GPU tests): it must issue stage k for every rank
 * before stage k+1 for any, so that no kernel ever spins on a flag that a later launch on the same GPU has to write. */
int fhe_b200_bfv_multiply_relin_sharded_stage(fhe_b200_shard* shard, int stage, const uint64_t* d_a, const uint64_t* d_b,
                                              const uint64_t* d_key_slice, uint64_t* d_out, uint32_t batch, void* stream);
/* synchronises `stream`; FHE_B200_ESTATE if a wait for a peer timed out since creation */
int fhe_b200_shard_check(fhe_b200_shard* shard, void* stream);
/* scheme constants for the compat layer and tests */
int fhe_b200_bfv_info(const fhe_b200_bfv* ctx, uint32_t* n, uint32_t* L, uint32_t* R, uint32_t* K, uint32_t* dnum,
                      uint64_t* t);
fhe_b200_plan* fhe_b200_bfv_plan(fhe_b200_bfv* ctx);
/* ---- wire format for keys / ciphertexts / plaintexts in HOST buffers (csrc/wire.cpp documents the 64-byte header) ---------
 * The reference declares no serialisation.  kind: 1 ciphertext, 2 public key, 3 secret key, 4 key-switching
 * key (relinearisation / Galois), 5 plaintext.  payload = uint64 [polys][limbs][n], little-endian, with FNV-1a checksums of
 * the payload and of the modulus chain; unpack verifies both, that the header's limb count equals n_moduli and that every word is reduced
 * modulo its limb (h_moduli may be NULL to skip the chain checks, h_words_out NULL to read the header only). */
size_t fhe_b200_wire_size(uint32_t n, uint32_t limbs, uint32_t polys);
int fhe_b200_wire_pack(uint32_t kind, uint32_t n, uint32_t limbs, uint32_t polys, int ntt_form, uint32_t galois_elt,
                       const uint64_t* h_moduli, const uint64_t* h_words, uint8_t* h_out);
int fhe_b200_wire_unpack(const uint8_t* h_in, size_t len, const uint64_t* h_moduli, uint32_t n_moduli, uint32_t* kind, uint32_t* n,
                         uint32_t* limbs, uint32_t* polys, int* ntt_form, uint32_t* galois_elt, uint64_t* h_words_out);
/* the Gaussian CDT the samplers use (<= 128 entries); returns the length */
int fhe_b200_gaussian_cdt(double sigma, uint64_t* h_cdt, uint32_t cap);

#ifdef __cplusplus
}
#endif
#endif
