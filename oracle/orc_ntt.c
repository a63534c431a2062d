/*
 * oracle/orc_ntt.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * CPU restatement of the reference's NTT / polynomial path:
 *   ring R_q = Z_q[x]/(x^N+1), a*b = INTT(NTT(a) (.) NTT(b))
 *       /root/reference/docs/ARCHITECTURE.md:225-241, include/polynomial.cuh:38-39
 *   butterflies (a,b)->(a+wb, a-wb) and (a,b)->(a+b,(a-b)w)
 *       /root/reference/include/ntt.cuh:147-167
 *   NTTEngine::forward / inverse / multiply
 *       /root/reference/src/ntt.cu:30-75
 *   element-wise add / sub / mul_scalar
 *       /root/reference/src/polynomial.cu:70-111
 *
 * PARITY PINNING.  The reference itself cannot be the oracle: its twiddle
 * table is 1,2,3,..., its root finder returns 3 and n^-1 is 1
 * (src/ntt.cu:86-119), so its NTT output is not a transform.  The oracle is
 * pinned instead by (i) the only known answers the reference's tests hold for
 * this path -- 12345+-67890 mod 100000 (tests/test_fhe.cu:34-56), the
 * forward->inverse identity at N=1024,q=12289,x_i=i+1 (tests/test_fhe.cu:68-116),
 * N=2048,q=40961 (tests/test_fhe.cu:126-167) -- and (ii) two independent
 * definitions kept in this file: the O(N^2) schoolbook negacyclic product and
 * the O(N^2) definitional transform X[k] = sum_j a_j psi^(j(2k+1)).
 *
 * Ordering contract (shared with the CUDA engine): forward takes natural
 * order and leaves X[k] at position bitrev(k); inverse takes that order back
 * to natural order.
 */
#include "orc_math.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------- primes and roots -------------------------------------------- */

/* deterministic Miller-Rabin for 64-bit n -- intent of reference is_prime, include/rns.cuh:146 */
int orc_is_prime(u64 n) {
    static const u64 bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return 0;
    for (int i = 0; i < 12; i++) { if (n % bases[i] == 0) return n == bases[i]; }
    u64 d = n - 1; int s = 0;
    while (!(d & 1)) { d >>= 1; s++; }
    for (int i = 0; i < 12; i++) {
        u64 x = orc_powmod(bases[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int r = 1; r < s; r++) { x = orc_mulmod(x, x, n); if (x == n - 1) { comp = 0; break; } }
        if (comp) return 0;
    }
    return 1;
}

/* out[k] = k-th largest prime p < 2^bits with p == 1 (mod 2^step_log2)
 * -- intent of reference generate_rns_primes / find_ntt_prime, include/rns.cuh:139-149 (SURVEY 8d prime chain) */
int orc_prime_chain(u32 bits, u32 step_log2, u32 count, u64 *out) {
    if (bits > 62 || bits < step_log2 + 1) return -1;
    u64 step = 1ULL << step_log2;
    u64 p = (1ULL << bits) - step + 1;
    u32 k = 0;
    while (k < count && p > step) {
        if (orc_is_prime(p)) out[k++] = p;
        p -= step;
    }
    return k == count ? 0 : -2;
}

/* psi rule: x = smallest integer >= 2 with x^((q-1)/2) == -1, psi = x^((q-1)/2N); order exactly 2N
 * -- intent of reference find_primitive_root / compute_root_of_unity, src/ntt.cu:100, include/ntt.cuh:102 */
u64 orc_find_psi(u64 q, u32 n) {
    if ((q - 1) % (2ULL * n)) return 0;
    for (u64 x = 2; x < q; x++) {
        if (orc_powmod(x, (q - 1) / 2, q) == q - 1) return orc_powmod(x, (q - 1) / (2ULL * n), q);
    }
    return 0;
}

/* fwd[k] = psi^bitrev(k), inv[k] = psi^-bitrev(k), k in [0,N)
 * -- intent of reference precompute_twiddle_factors, src/ntt.cu:77-108 */
void orc_ntt_tables(u64 q, u32 n, u64 psi, u64 *fwd, u64 *inv) {
    u32 lg = orc_log2(n);
    u64 ipsi = orc_invmod_prime(psi, q);
    u64 p = 1, ip = 1;
    for (u32 e = 0; e < n; e++) {
        u32 k = orc_bitrev(e, lg);
        fwd[k] = p; inv[k] = ip;
        p = orc_mulmod(p, psi, q); ip = orc_mulmod(ip, ipsi, q);
    }
}

/* ---------- definitional references (O(N^2)) ---------------------------- */

/* X[k] = sum_j a_j psi^(j(2k+1)) mod q, natural order in k */
void orc_negacyclic_dft_def(u64 *out, const u64 *a, u32 n, u64 q, u64 psi) {
    for (u32 k = 0; k < n; k++) {
        u64 w = orc_powmod(psi, 2ULL * k + 1, q), wj = 1, acc = 0;
        for (u32 j = 0; j < n; j++) {
            acc = orc_addmod(acc, orc_mulmod(a[j] % q, wj, q), q);
            wj = orc_mulmod(wj, w, q);
        }
        out[k] = acc;
    }
}

/* c = a*b mod (x^N+1, q), schoolbook -- the C1 ground truth */
void orc_schoolbook_negacyclic(u64 *out, const u64 *a, const u64 *b, u32 n, u64 q) {
    #pragma omp parallel for schedule(static)
    for (u32 k = 0; k < n; k++) {
        u64 pos = 0, neg = 0;
        for (u32 i = 0; i <= k; i++) pos = orc_addmod(pos, orc_mulmod(a[i], b[k - i], q), q);
        for (u32 i = k + 1; i < n; i++) neg = orc_addmod(neg, orc_mulmod(a[i], b[n + k - i], q), q);
        out[k] = orc_submod(pos, neg, q);
    }
}

/* ---------- O(N log N) transforms --------------------------------------- */

/* in place; natural -> bit-reversed.  CT butterflies, ntt.cuh:147-156 */
void orc_ntt_forward_tab(u64 *a, u32 n, u64 q, const u64 *fwd) {
    u32 t = n >> 1;
    for (u32 m = 1; m < n; m <<= 1, t >>= 1) {
        for (u32 i = 0; i < m; i++) {
            u64 w = fwd[m + i];
            u64 *x = a + 2 * (size_t)i * t, *y = x + t;
            for (u32 j = 0; j < t; j++) {
                u64 u = x[j], v = orc_mulmod(y[j], w, q);
                x[j] = orc_addmod(u, v, q);
                y[j] = orc_submod(u, v, q);
            }
        }
    }
}

/* in place; bit-reversed -> natural, includes the 1/N scaling.  GS butterflies, ntt.cuh:158-167 */
void orc_ntt_inverse_tab(u64 *a, u32 n, u64 q, const u64 *inv) {
    u32 t = 1;
    for (u32 m = n >> 1; m >= 1; m >>= 1, t <<= 1) {
        for (u32 i = 0; i < m; i++) {
            u64 w = inv[m + i];
            u64 *x = a + 2 * (size_t)i * t, *y = x + t;
            for (u32 j = 0; j < t; j++) {
                u64 u = x[j], v = y[j];
                x[j] = orc_addmod(u, v, q);
                y[j] = orc_mulmod(orc_submod(u, v, q), w, q);
            }
        }
    }
    u64 ninv = orc_invmod_prime(n % q, q);
    for (u32 j = 0; j < n; j++) a[j] = orc_mulmod(a[j], ninv, q);
}

int orc_ntt_forward(u64 *a, u32 n, u64 q) {
    u64 psi = orc_find_psi(q, n);
    if (!psi) return -1;
    u64 *f = (u64 *)malloc(2 * (size_t)n * sizeof(u64));
    orc_ntt_tables(q, n, psi, f, f + n);
    orc_ntt_forward_tab(a, n, q, f);
    free(f);
    return 0;
}
int orc_ntt_inverse(u64 *a, u32 n, u64 q) {
    u64 psi = orc_find_psi(q, n);
    if (!psi) return -1;
    u64 *f = (u64 *)malloc(2 * (size_t)n * sizeof(u64));
    orc_ntt_tables(q, n, psi, f, f + n);
    orc_ntt_inverse_tab(a, n, q, f + n);
    free(f);
    return 0;
}

/* NTTEngine::multiply, src/ntt.cu:49-75: out = a*b in R_q */
int orc_negacyclic_mul_ntt(u64 *out, const u64 *a, const u64 *b, u32 n, u64 q) {
    u64 psi = orc_find_psi(q, n);
    if (!psi) return -1;
    u64 *f = (u64 *)malloc(4 * (size_t)n * sizeof(u64));
    u64 *ta = f + 2 * (size_t)n, *tb = ta + n;
    orc_ntt_tables(q, n, psi, f, f + n);
    memcpy(ta, a, n * sizeof(u64)); memcpy(tb, b, n * sizeof(u64));
    orc_ntt_forward_tab(ta, n, q, f); orc_ntt_forward_tab(tb, n, q, f);
    for (u32 i = 0; i < n; i++) ta[i] = orc_mulmod(ta[i], tb[i], q);
    orc_ntt_inverse_tab(ta, n, q, f + n);
    memcpy(out, ta, n * sizeof(u64));
    free(f);
    return 0;
}

/* ---------- element-wise, src/polynomial.cu:70-111, src/rns.cu:143-180 --- */
/* layout [batch][limb][N]; limb l uses moduli[l] */
void orc_poly_add(u64 *o, const u64 *a, const u64 *b, u32 n, const u64 *moduli, u32 limbs, u32 batch) {
    for (size_t p = 0; p < (size_t)batch * limbs; p++) { u64 q = moduli[p % limbs];
        for (u32 j = 0; j < n; j++) o[p * n + j] = orc_addmod(a[p * n + j], b[p * n + j], q); }
}
void orc_poly_sub(u64 *o, const u64 *a, const u64 *b, u32 n, const u64 *moduli, u32 limbs, u32 batch) {
    for (size_t p = 0; p < (size_t)batch * limbs; p++) { u64 q = moduli[p % limbs];
        for (u32 j = 0; j < n; j++) o[p * n + j] = orc_submod(a[p * n + j], b[p * n + j], q); }
}
void orc_poly_mul(u64 *o, const u64 *a, const u64 *b, u32 n, const u64 *moduli, u32 limbs, u32 batch) {
    for (size_t p = 0; p < (size_t)batch * limbs; p++) { u64 q = moduli[p % limbs];
        for (u32 j = 0; j < n; j++) o[p * n + j] = orc_mulmod(a[p * n + j], b[p * n + j], q); }
}
/* o = acc + a*b */
void orc_poly_mac(u64 *o, const u64 *acc, const u64 *a, const u64 *b, u32 n, const u64 *moduli, u32 limbs, u32 batch) {
    for (size_t p = 0; p < (size_t)batch * limbs; p++) { u64 q = moduli[p % limbs];
        for (u32 j = 0; j < n; j++) o[p * n + j] = orc_addmod(acc[p * n + j], orc_mulmod(a[p * n + j], b[p * n + j], q), q); }
}
/* scalars[l] is the per-limb residue of the scalar */
void orc_poly_mul_scalar(u64 *o, const u64 *a, const u64 *scalars, u32 n, const u64 *moduli, u32 limbs, u32 batch) {
    for (size_t p = 0; p < (size_t)batch * limbs; p++) { u64 q = moduli[p % limbs], s = scalars[p % limbs] % q;
        for (u32 j = 0; j < n; j++) o[p * n + j] = orc_mulmod(a[p * n + j], s, q); }
}

/* ---------- batched RNS transforms: CPU baseline (Shoup/Harvey, OpenMP) -- */
/* RNS_NTTEngine::forward_rns / inverse_rns, src/ntt.cu:158-171, over [batch][limbs][N].
 * Same results as orc_ntt_forward_tab; written with Shoup mulmod so the
 * reported CPU baseline is a fair one.  tables: [limbs][4][N] = fwd, fwd_shoup, inv, inv_shoup. */
static inline u64 shoup_mul(u64 x, u64 w, u64 ws, u64 q) {
    u64 h = (u64)(((u128)x * ws) >> 64);
    u64 r = x * w - h * q;
    return r >= q ? r - q : r;
}
void orc_rns_tables(u64 *tables, const u64 *moduli, u32 limbs, u32 n) {
    for (u32 l = 0; l < limbs; l++) {
        u64 q = moduli[l], *T = tables + (size_t)l * 4 * n;
        u64 psi = orc_find_psi(q, n);
        orc_ntt_tables(q, n, psi, T, T + 2 * (size_t)n);
        for (u32 k = 0; k < n; k++) {
            T[(size_t)n + k] = (u64)((((u128)T[k]) << 64) / q);
            T[3 * (size_t)n + k] = (u64)((((u128)T[2 * (size_t)n + k]) << 64) / q);
        }
    }
}
/* Harvey's lazy butterflies (values kept in [0, 4q) forward / [0, 2q) inverse, one full reduction at the end): the form a tuned
 * CPU library uses, so that the reported CPU baseline is not handicapped by per-butterfly conditional corrections.  q < 2^61. */
static inline u64 shoup_lazy(u64 x, u64 w, u64 ws, u64 q) {          /* x*w mod q + {0, q}, for any 64-bit x */
    u64 h = (u64)(((u128)x * ws) >> 64);
    return x * w - h * q;
}
static void fwd_shoup(u64 *a, u32 n, u64 q, const u64 *w, const u64 *ws) {
    const u64 q2 = 2 * q;
    u32 t = n >> 1;
    for (u32 m = 1; m < n; m <<= 1, t >>= 1)
        for (u32 i = 0; i < m; i++) {
            u64 W = w[m + i], Ws = ws[m + i];
            u64 *x = a + 2 * (size_t)i * t, *y = x + t;
            for (u32 j = 0; j < t; j++) {
                u64 u = x[j]; u = u >= q2 ? u - q2 : u;
                u64 v = shoup_lazy(y[j], W, Ws, q);
                x[j] = u + v;
                y[j] = u + q2 - v;
            }
        }
    for (u32 j = 0; j < n; j++) { u64 u = a[j]; u = u >= q2 ? u - q2 : u; a[j] = u >= q ? u - q : u; }
}
static void inv_shoup(u64 *a, u32 n, u64 q, const u64 *w, const u64 *ws) {
    const u64 q2 = 2 * q;
    u32 t = 1;
    for (u32 m = n >> 1; m >= 1; m >>= 1, t <<= 1)
        for (u32 i = 0; i < m; i++) {
            u64 W = w[m + i], Ws = ws[m + i];
            u64 *x = a + 2 * (size_t)i * t, *y = x + t;
            for (u32 j = 0; j < t; j++) {
                u64 u = x[j], v = y[j];
                u64 s = u + v; x[j] = s >= q2 ? s - q2 : s;
                y[j] = shoup_lazy(u + q2 - v, W, Ws, q);
            }
        }
    u64 ninv = orc_invmod_prime(n % q, q), ninvs = (u64)((((u128)ninv) << 64) / q);
    for (u32 j = 0; j < n; j++) { u64 r = shoup_lazy(a[j], ninv, ninvs, q); a[j] = r >= q ? r - q : r; }
}
/* returns the number of OpenMP threads used */
int orc_rns_ntt_batch(u64 *data, const u64 *tables, const u64 *moduli, u32 limbs, u32 n, u32 batch, int inverse, int threads) {
    int used = 1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    used = omp_get_max_threads();
#endif
    long total = (long)batch * limbs;
    #pragma omp parallel for schedule(dynamic, 1)
    for (long p = 0; p < total; p++) {
        u32 l = (u32)(p % limbs);
        const u64 *T = tables + (size_t)l * 4 * n;
        if (inverse) inv_shoup(data + (size_t)p * n, n, moduli[l], T + 2 * (size_t)n, T + 3 * (size_t)n);
        else fwd_shoup(data + (size_t)p * n, n, moduli[l], T, T + (size_t)n);
    }
    return used;
}
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* reference unit test values: batch_mod_add/sub, tests/test_fhe.cu:34-56 */
u64 orc_add_mod(u64 a, u64 b, u64 q) { return orc_addmod(a % q, b % q, q); }
u64 orc_sub_mod(u64 a, u64 b, u64 q) { return orc_submod(a % q, b % q, q); }
u64 orc_mul_mod(u64 a, u64 b, u64 q) { return orc_mulmod(a % q, b % q, q); }
u64 orc_pow_mod(u64 a, u64 e, u64 q) { return orc_powmod(a, e, q); }
u64 orc_inv_mod(u64 a, u64 q) { return orc_invmod_prime(a, q); }
u32 orc_chacha_key[8];
int orc_chacha_on = 0;
/* key32 = NULL: back to the splitmix generator */
void orc_set_rng_key(const unsigned char *key32) {
    orc_chacha_on = key32 != NULL;
    for (int i = 0; i < 8; i++)
        orc_chacha_key[i] = key32 ? ((u32)key32[4 * i] | ((u32)key32[4 * i + 1] << 8) | ((u32)key32[4 * i + 2] << 16) | ((u32)key32[4 * i + 3] << 24)) : 0;
}
/* RFC 8439 block function on explicit inputs (known-answer test) */
void orc_chacha20_block_raw(const u32 *key, u32 counter, const u32 *nonce, u32 *out) { orc_chacha20_block(key, counter, nonce, out); }
u64 orc_rng(u64 seed, u64 stream, u64 idx) { return orc_rng64(seed, stream, idx); }
u64 orc_rng_key_of(u64 seed, u64 stream) { return orc_rng_key(seed, stream); }
u64 orc_item_seed_of(u64 seed, u64 item) { return orc_item_seed(seed, item); }
