/*
 * oracle/orc_bfv.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * CPU restatement of the reference's BFV path under fhe::FHEContext:
 *   keygen          /root/reference/src/fhe.cu:54-74    pk = (e - a*s, a), s ternary
 *   relinkey_gen    /root/reference/src/fhe.cu:76-111   key_j = (-a_j s + e_j + B_j s^2, a_j)
 *   encode/decode   /root/reference/src/fhe.cu:113-136  coefficient encoding
 *   encrypt         /root/reference/src/fhe.cu:138-169  c0 = pk0 u + e1 + Delta m, c1 = pk1 u + e2
 *   decrypt         /root/reference/src/fhe.cu:171-185  m = round(t/q (c0 + c1 s)) mod t
 *   add             /root/reference/src/fhe.cu:187-197
 *   multiply        /root/reference/src/fhe.cu:199-224  tensor (c0,c1,c2) [+ the t/q scaling the docs require,
 *                                                       docs/ARCHITECTURE.md:306-317]
 *   relinearize     /root/reference/src/fhe.cu:226-235  (stub there) -> docs/ARCHITECTURE.md:319-326
 *
 * The reference hard-codes q = 2^60 and never scales or relinearises, so its
 * outputs are not a BFV result; what is restated is the documented scheme in
 * RNS form:  Q = q_0..q_{L-1};  auxiliary basis R = p_0..p_{R-1} for the
 * tensor product;  hybrid (dnum-digit) key switching with the first K
 * auxiliary primes as special modulus P (BASELINE.json config 4).
 * PARITY PINNING: tests/test_oracle_bfv.py re-derives every step with Python
 * big integers (exact CRT, exact rational rounding) at small N.
 *
 * Polynomials are [limb][N] uint64.  Ciphertexts and plaintexts are in
 * coefficient form (the reference's is_ntt_form = false); keys are kept in
 * NTT form (bit-reversed order).
 */
#include "orc_math.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct orc_lincomb orc_lincomb;
orc_lincomb *orc_lc_make_conv(const u64 *src, u32 S, const u64 *dst, u32 T);
orc_lincomb *orc_lc_make_scale(const u64 *qs, u32 L, const u64 *ps, u32 R, u64 t, const u64 *targets, u32 T, int with_extra);
void orc_lc_apply(const orc_lincomb *lc, u64 *out, const u64 *in, const u64 *extra, u32 n);
void orc_lc_free(orc_lincomb *lc);
u64 orc_inv_general(u64 a, u64 m);
void orc_ntt_tables(u64 q, u32 n, u64 psi, u64 *fwd, u64 *inv);
u64 orc_find_psi(u64 q, u32 n);
void orc_ntt_forward_tab(u64 *a, u32 n, u64 q, const u64 *fwd);
void orc_ntt_inverse_tab(u64 *a, u32 n, u64 q, const u64 *inv);

typedef struct {
    u32 n, L, R, K, dnum, alpha;
    u64 t;
    u64 *primes;        /* [L+R]: Q then aux; special P = aux[0..K) */
    u64 *fwd, *inv;     /* [L+R][n] */
    u64 *delta;         /* [L]: floor(Q/t) mod q_i */
    u64 *p_mod_q;       /* [L]: P mod q_i */
    u64 *pinv_mod_q;    /* [L]: P^-1 mod q_i */
    orc_lincomb *q2r, *scale, *r2q, *moddown, *dec;
    orc_lincomb **modup;   /* [dnum] */
    u32 **modup_targets;   /* [dnum][L+K-alpha] indices into primes */
} orc_bfv;

static void fwd_limb(const orc_bfv *c, u64 *a, u32 pi) { orc_ntt_forward_tab(a, c->n, c->primes[pi], c->fwd + (size_t)pi * c->n); }
static void inv_limb(const orc_bfv *c, u64 *a, u32 pi) { orc_ntt_inverse_tab(a, c->n, c->primes[pi], c->inv + (size_t)pi * c->n); }

orc_bfv *orc_bfv_create(u32 n, u32 L, u32 R, u32 K, u32 dnum, u64 t, const u64 *primes) {
    if (L % dnum || K > R || L + R > 120) return NULL;
    orc_bfv *c = (orc_bfv *)calloc(1, sizeof(*c));
    c->n = n; c->L = L; c->R = R; c->K = K; c->dnum = dnum; c->alpha = L / dnum; c->t = t;
    u32 A = L + R;
    c->primes = (u64 *)malloc(A * sizeof(u64)); memcpy(c->primes, primes, A * sizeof(u64));
    c->fwd = (u64 *)malloc((size_t)A * n * sizeof(u64)); c->inv = (u64 *)malloc((size_t)A * n * sizeof(u64));
    for (u32 i = 0; i < A; i++) {
        u64 psi = orc_find_psi(primes[i], n);
        if (!psi) { free(c); return NULL; }
        orc_ntt_tables(primes[i], n, psi, c->fwd + (size_t)i * n, c->inv + (size_t)i * n);
    }
    const u64 *Q = primes, *P = primes + L;
    c->delta = (u64 *)malloc(L * sizeof(u64)); c->p_mod_q = (u64 *)malloc(L * sizeof(u64)); c->pinv_mod_q = (u64 *)malloc(L * sizeof(u64));
    u64 q_mod_t = 1 % t;
    for (u32 i = 0; i < L; i++) q_mod_t = orc_mulmod(q_mod_t, Q[i] % t, t);
    for (u32 i = 0; i < L; i++) {
        u64 qi = Q[i];
        /* floor(Q/t) = (Q - (Q mod t))/t  ==>  mod q_i: -(Q mod t) * t^-1 */
        c->delta[i] = orc_mulmod(orc_negmod(q_mod_t % qi, qi), orc_inv_general(t % qi, qi), qi);
        u64 pm = 1;
        for (u32 k = 0; k < K; k++) pm = orc_mulmod(pm, P[k] % qi, qi);
        c->p_mod_q[i] = pm; c->pinv_mod_q[i] = orc_inv_general(pm, qi);
    }
    c->q2r = orc_lc_make_conv(Q, L, P, R);
    c->scale = orc_lc_make_scale(Q, L, P, R, t, P, R, 1);
    c->r2q = orc_lc_make_conv(P, R, Q, L);
    c->moddown = orc_lc_make_conv(P, K, Q, L);
    c->dec = orc_lc_make_scale(Q, L, NULL, 0, t, &c->t, 1, 0);
    c->modup = (orc_lincomb **)calloc(dnum, sizeof(void *));
    c->modup_targets = (u32 **)calloc(dnum, sizeof(void *));
    for (u32 d = 0; d < dnum; d++) {
        u32 nt = L + K - c->alpha, k = 0;
        u64 *tm = (u64 *)malloc(nt * sizeof(u64));
        c->modup_targets[d] = (u32 *)malloc(nt * sizeof(u32));
        for (u32 i = 0; i < L + K; i++) {
            if (i >= d * c->alpha && i < (d + 1) * c->alpha) continue;
            c->modup_targets[d][k] = i; tm[k] = primes[i]; k++;
        }
        c->modup[d] = orc_lc_make_conv(Q + d * c->alpha, c->alpha, tm, nt);
        free(tm);
    }
    return c;
}
void orc_bfv_destroy(orc_bfv *c) {
    if (!c) return;
    for (u32 d = 0; d < c->dnum; d++) { orc_lc_free(c->modup[d]); free(c->modup_targets[d]); }
    free(c->modup); free(c->modup_targets);
    orc_lc_free(c->q2r); orc_lc_free(c->scale); orc_lc_free(c->r2q); orc_lc_free(c->moddown); orc_lc_free(c->dec);
    free(c->primes); free(c->fwd); free(c->inv); free(c->delta); free(c->p_mod_q); free(c->pinv_mod_q); free(c);
}
void orc_bfv_get_consts(const orc_bfv *c, u64 *delta, u64 *p_mod_q, u64 *pinv_mod_q) {
    memcpy(delta, c->delta, c->L * sizeof(u64)); memcpy(p_mod_q, c->p_mod_q, c->L * sizeof(u64));
    memcpy(pinv_mod_q, c->pinv_mod_q, c->L * sizeof(u64));
}

/* ---------- samplers (specification shared with the CUDA kernels) -------- */
/* intent of sample_ternary_kernel / sample_gaussian_kernel / sample_uniform_kernel,
 * /root/reference/include/polynomial.cuh:112-135 (placeholders in src/polynomial.cu:113-143) */

/* cdt[k] = floor(2^63 * P(|X| <= k)), X ~ discrete Gaussian(sigma), tail cut at ceil(6 sigma); returns length */
u32 orc_gaussian_cdt(double sigma, u64 *cdt, u32 cap) {
    u32 tail = (u32)ceil(6.0 * sigma);
    if (tail + 1 > cap) tail = cap - 1;
    double s = 1.0;
    for (u32 x = 1; x <= tail; x++) s += 2.0 * exp(-((double)x * (double)x) / (2.0 * sigma * sigma));
    double cum = 1.0;
    for (u32 k = 0; k <= tail; k++) {
        if (k) cum += 2.0 * exp(-((double)k * (double)k) / (2.0 * sigma * sigma));
        double f = cum / s;
        cdt[k] = f >= 1.0 ? (1ULL << 63) : (u64)(f * 9223372036854775808.0);
    }
    cdt[tail] = 1ULL << 63;
    return tail + 1;
}
static inline int64_t gauss_from(u64 r, const u64 *cdt, u32 len) {
    u64 u = r >> 1; int64_t mag = 0;
    for (u32 k = 0; k + 1 < len; k++) mag += (u >= cdt[k]);
    return (r & 1) ? -mag : mag;
}
static inline int64_t ternary_from(u64 r, u32 thr) {
    if ((u32)(r >> 32) >= thr) return 0;
    return (r & 1) ? -1 : 1;
}
void orc_sample_ternary(int64_t *out, u32 n, u64 seed, u64 stream, u32 thr) {
    for (u32 j = 0; j < n; j++) out[j] = ternary_from(orc_rng64(seed, stream, j), thr);
}
void orc_sample_gaussian(int64_t *out, u32 n, u64 seed, u64 stream, const u64 *cdt, u32 len) {
    for (u32 j = 0; j < n; j++) out[j] = gauss_from(orc_rng64(seed, stream, j), cdt, len);
}
/* exactly hw non-zero coefficients: partial Fisher-Yates on stream, signs on stream+1 */
void orc_sample_ternary_hw(int64_t *out, u32 n, u64 seed, u64 stream, u32 hw) {
    u32 *perm = (u32 *)malloc(n * sizeof(u32));
    for (u32 j = 0; j < n; j++) { perm[j] = j; out[j] = 0; }
    if (hw > n) hw = n;
    for (u32 i = 0; i < hw; i++) {
        u32 j = i + (u32)(orc_rng64(seed, stream, i) % (n - i));
        u32 tmp = perm[i]; perm[i] = perm[j]; perm[j] = tmp;
        out[perm[i]] = (orc_rng64(seed, stream + 1, i) & 1) ? -1 : 1;
    }
    free(perm);
}
/* uniform in [0,q): (r0*2^64 + r1) mod q */
void orc_sample_uniform(u64 *out, u32 n, u64 q, u64 seed, u64 stream) {
    for (u32 j = 0; j < n; j++) {
        u128 v = ((u128)orc_rng64(seed, stream, 2ULL * j) << 64) | orc_rng64(seed, stream, 2ULL * j + 1);
        out[j] = (u64)(v % q);
    }
}
static void small_to_limb(u64 *out, const int64_t *s, u32 n, u64 q) {
    for (u32 j = 0; j < n; j++) out[j] = s[j] >= 0 ? (u64)s[j] % q : q - ((u64)(-s[j]) % q);
}

/* stream ids (part of the specification) */
enum { ST_SK = 0, ST_SK_SIGN = 1, ST_PK_E = 2, ST_PK_A = 16,
       ST_ENC_U = 0, ST_ENC_E1 = 1, ST_ENC_E2 = 2,
       ST_RLK_BASE = 1024, ST_RLK_E = 512 };

/* ---------- keys --------------------------------------------------------- */
/* sk_ntt: [L+R][n] NTT form over every prime;  hw == 0 -> ternary with P(nonzero)=thr/2^32 (reference: 0.5, src/fhe.cu:254-256) */
void orc_bfv_secret_keygen(const orc_bfv *c, u64 seed, u32 hw, u32 thr, int64_t *s_small, u64 *sk_ntt) {
    u32 n = c->n, A = c->L + c->R;
    if (hw) orc_sample_ternary_hw(s_small, n, seed, ST_SK, hw);
    else orc_sample_ternary(s_small, n, seed, ST_SK, thr);
    for (u32 i = 0; i < A; i++) { small_to_limb(sk_ntt + (size_t)i * n, s_small, n, c->primes[i]); fwd_limb(c, sk_ntt + (size_t)i * n, i); }
}
/* pk: [2][L][n] NTT form;  pk0 = e - a*s, pk1 = a   (src/fhe.cu:54-74) */
void orc_bfv_public_keygen(const orc_bfv *c, u64 seed, const u64 *sk_ntt, const u64 *cdt, u32 cdt_len, u64 *pk) {
    u32 n = c->n, L = c->L;
    int64_t *e = (int64_t *)malloc(n * sizeof(int64_t));
    orc_sample_gaussian(e, n, seed, ST_PK_E, cdt, cdt_len);
    for (u32 i = 0; i < L; i++) {
        u64 q = c->primes[i];
        u64 *pk0 = pk + (size_t)i * n, *pk1 = pk + (size_t)(L + i) * n;
        orc_sample_uniform(pk1, n, q, seed, ST_PK_A + i);
        small_to_limb(pk0, e, n, q); fwd_limb(c, pk0, i);
        for (u32 j = 0; j < n; j++) pk0[j] = orc_submod(pk0[j], orc_mulmod(pk1[j], sk_ntt[(size_t)i * n + j], q), q);
    }
    free(e);
}
/* rlk: [dnum][2][L+K][n] NTT form;  b_d = -a_d s + e_d + P*B_d*s^2, B_d = 1 mod Q_d, 0 mod Q/Q_d   (src/fhe.cu:76-111) */
void orc_bfv_relin_keygen(const orc_bfv *c, u64 seed, const u64 *sk_ntt, const u64 *cdt, u32 cdt_len, u64 *rlk) {
    u32 n = c->n, L = c->L, K = c->K, W = L + K;
    int64_t *e = (int64_t *)malloc(n * sizeof(int64_t));
    for (u32 d = 0; d < c->dnum; d++) {
        orc_sample_gaussian(e, n, seed, ST_RLK_BASE * (d + 1) + ST_RLK_E, cdt, cdt_len);
        for (u32 i = 0; i < W; i++) {
            u64 q = c->primes[i];
            u64 *b = rlk + ((size_t)(d * 2 + 0) * W + i) * n, *a = rlk + ((size_t)(d * 2 + 1) * W + i) * n;
            const u64 *s = sk_ntt + (size_t)i * n;
            orc_sample_uniform(a, n, q, seed, ST_RLK_BASE * (d + 1) + i);
            small_to_limb(b, e, n, q); fwd_limb(c, b, i);
            int in_group = (i < L) && (i / c->alpha == d);
            u64 f = in_group ? c->p_mod_q[i] : 0;
            for (u32 j = 0; j < n; j++) {
                u64 v = orc_submod(b[j], orc_mulmod(a[j], s[j], q), q);
                if (f) v = orc_addmod(v, orc_mulmod(f, orc_mulmod(s[j], s[j], q), q), q);
                b[j] = v;
            }
        }
    }
    free(e);
}

/* ---------- encrypt / decrypt / add -------------------------------------- */
/* pt: [n] coefficients mod t;  ct: [2][L][n] coefficient form   (src/fhe.cu:138-169) */
void orc_bfv_encrypt(const orc_bfv *c, u64 seed, const u64 *pt, const u64 *pk, u32 thr, const u64 *cdt, u32 cdt_len, u64 *ct) {
    u32 n = c->n, L = c->L;
    int64_t *u = (int64_t *)malloc(3 * (size_t)n * sizeof(int64_t)), *e1 = u + n, *e2 = e1 + n;
    u64 *un = (u64 *)malloc(n * sizeof(u64));
    orc_sample_ternary(u, n, seed, ST_ENC_U, thr);
    orc_sample_gaussian(e1, n, seed, ST_ENC_E1, cdt, cdt_len);
    orc_sample_gaussian(e2, n, seed, ST_ENC_E2, cdt, cdt_len);
    for (u32 i = 0; i < L; i++) {
        u64 q = c->primes[i];
        u64 *c0 = ct + (size_t)i * n, *c1 = ct + (size_t)(L + i) * n;
        small_to_limb(un, u, n, q); fwd_limb(c, un, i);
        for (u32 j = 0; j < n; j++) { c0[j] = orc_mulmod(pk[(size_t)i * n + j], un[j], q); c1[j] = orc_mulmod(pk[(size_t)(L + i) * n + j], un[j], q); }
        inv_limb(c, c0, i); inv_limb(c, c1, i);
        for (u32 j = 0; j < n; j++) {
            u64 ev1 = e1[j] >= 0 ? (u64)e1[j] : q - (u64)(-e1[j]);
            u64 ev2 = e2[j] >= 0 ? (u64)e2[j] : q - (u64)(-e2[j]);
            u64 dm = orc_mulmod(c->delta[i], pt[j] % q, q);
            c0[j] = orc_addmod(orc_addmod(c0[j], ev1, q), dm, q);
            c1[j] = orc_addmod(c1[j], ev2, q);
        }
    }
    free(u); free(un);
}
/* (src/fhe.cu:171-185)  pt[j] = round(t/Q * [c0 + c1 s]_Q) mod t */
void orc_bfv_decrypt(const orc_bfv *c, const u64 *ct, const u64 *sk_ntt, u64 *pt) {
    u32 n = c->n, L = c->L;
    u64 *x = (u64 *)malloc((size_t)L * n * sizeof(u64));
    for (u32 i = 0; i < L; i++) {
        u64 q = c->primes[i]; u64 *xi = x + (size_t)i * n;
        memcpy(xi, ct + (size_t)(L + i) * n, n * sizeof(u64));
        fwd_limb(c, xi, i);
        for (u32 j = 0; j < n; j++) xi[j] = orc_mulmod(xi[j], sk_ntt[(size_t)i * n + j], q);
        inv_limb(c, xi, i);
        for (u32 j = 0; j < n; j++) xi[j] = orc_addmod(xi[j], ct[(size_t)i * n + j], q);
    }
    orc_lc_apply(c->dec, pt, x, NULL, n);
    free(x);
}
/* (src/fhe.cu:187-197) */
void orc_bfv_add(const orc_bfv *c, const u64 *a, const u64 *b, u64 *out) {
    u32 n = c->n, L = c->L;
    for (u32 p = 0; p < 2 * L; p++) { u64 q = c->primes[p % L];
        for (u32 j = 0; j < n; j++) out[(size_t)p * n + j] = orc_addmod(a[(size_t)p * n + j], b[(size_t)p * n + j], q); }
}

/* Hybrid key switching (docs/ARCHITECTURE.md:319-326; the reference's relinearize, src/fhe.cu:226-235, is a stub):
 *   out_p = add_p + ModDown( sum_digits ModUp(x digit) * key[digit][p] ),  p = 0, 1;  add_p may be NULL.
 * x, add_p: [L][n] coefficient form; key: [dnum][2][L+K][n] NTT form; out: [2][L][n] coefficient form. */
static void key_switch(const orc_bfv *c, const u64 *x, const u64 *key, const u64 *add0, const u64 *add1, u64 *out) {
    u32 n = c->n, L = c->L, K = c->K, W = L + K, alpha = c->alpha;
    u64 *acc = (u64 *)calloc(2 * (size_t)W * n, sizeof(u64));
    u64 *dig = (u64 *)malloc((size_t)W * n * sizeof(u64));
    u64 *conv = (u64 *)malloc((size_t)W * n * sizeof(u64));
    for (u32 dg = 0; dg < c->dnum; dg++) {
        orc_lc_apply(c->modup[dg], conv, x + (size_t)dg * alpha * n, NULL, n);    /* ModUp */
        u32 nt = W - alpha;
        for (u32 k = 0; k < nt; k++) memcpy(dig + (size_t)c->modup_targets[dg][k] * n, conv + (size_t)k * n, n * sizeof(u64));
        for (u32 i = dg * alpha; i < (dg + 1) * alpha; i++) memcpy(dig + (size_t)i * n, x + (size_t)i * n, n * sizeof(u64));
        for (u32 i = 0; i < W; i++) {
            u64 q = c->primes[i]; u64 *v = dig + (size_t)i * n;
            fwd_limb(c, v, i);
            const u64 *kb = key + ((size_t)(dg * 2 + 0) * W + i) * n, *ka = key + ((size_t)(dg * 2 + 1) * W + i) * n;
            u64 *s0 = acc + (size_t)i * n, *s1 = acc + (size_t)(W + i) * n;
            for (u32 j = 0; j < n; j++) {
                s0[j] = orc_addmod(s0[j], orc_mulmod(v[j], kb[j], q), q);
                s1[j] = orc_addmod(s1[j], orc_mulmod(v[j], ka[j], q), q);
            }
        }
    }
    for (int p = 0; p < 2; p++) {
        u64 *s = acc + (size_t)p * W * n;
        const u64 *add = p ? add1 : add0;
        for (u32 i = 0; i < W; i++) inv_limb(c, s + (size_t)i * n, i);
        orc_lc_apply(c->moddown, conv, s + (size_t)L * n, NULL, n);                /* [s]_P -> Q (centred) */
        for (u32 i = 0; i < L; i++) {
            u64 q = c->primes[i];
            for (u32 j = 0; j < n; j++) {
                u64 v = orc_mulmod(orc_submod(s[(size_t)i * n + j], conv[(size_t)i * n + j], q), c->pinv_mod_q[i], q);
                out[((size_t)p * L + i) * n + j] = add ? orc_addmod(add[(size_t)i * n + j], v, q) : v;
            }
        }
    }
    free(acc); free(dig); free(conv);
}

/* ---------- multiply, relinearize ----------------------------------------- */
/* The reference's multiply builds the three tensor components and then calls relinearize (src/fhe.cu:198-224; its
 * relinearize, :226-235, is a stub and its tensor is never scaled by t/Q).  Here:
 *   sc  : [3][L][n]  round(t/Q * tensor) in basis Q, coefficient form -- the 3-component ciphertext */
static void tensor_scale(const orc_bfv *c, const u64 *cta, const u64 *ctb, u64 *sc) {
    u32 n = c->n, L = c->L, R = c->R, A = L + R;
    size_t pn = (size_t)A * n;
    u64 *ext = (u64 *)malloc(4 * pn * sizeof(u64));       /* a0,a1,b0,b1 over Q u R */
    u64 *d = (u64 *)malloc(3 * pn * sizeof(u64));         /* tensor over Q u R */
    const u64 *in[4] = {cta, cta + (size_t)L * n, ctb, ctb + (size_t)L * n};
    for (int p = 0; p < 4; p++) {
        u64 *e = ext + p * pn;
        memcpy(e, in[p], (size_t)L * n * sizeof(u64));
        orc_lc_apply(c->q2r, e + (size_t)L * n, in[p], NULL, n);          /* 1. exact Q -> R */
        for (u32 i = 0; i < A; i++) fwd_limb(c, e + (size_t)i * n, i);    /* 2. NTT          */
    }
    for (u32 i = 0; i < A; i++) {                                          /* 3. tensor       */
        u64 q = c->primes[i];
        const u64 *a0 = ext + (size_t)i * n, *a1 = ext + pn + (size_t)i * n, *b0 = ext + 2 * pn + (size_t)i * n, *b1 = ext + 3 * pn + (size_t)i * n;
        u64 *d0 = d + (size_t)i * n, *d1 = d + pn + (size_t)i * n, *d2 = d + 2 * pn + (size_t)i * n;
        for (u32 j = 0; j < n; j++) {
            d0[j] = orc_mulmod(a0[j], b0[j], q);
            d1[j] = orc_addmod(orc_mulmod(a0[j], b1[j], q), orc_mulmod(a1[j], b0[j], q), q);
            d2[j] = orc_mulmod(a1[j], b1[j], q);
        }
        inv_limb(c, d0, i); inv_limb(c, d1, i); inv_limb(c, d2, i);       /* 4. INTT         */
    }
    u64 *tmpR = (u64 *)malloc((size_t)R * n * sizeof(u64));
    for (int p = 0; p < 3; p++) {
        orc_lc_apply(c->scale, tmpR, d + p * pn, d + p * pn + (size_t)L * n, n);   /* 5. round(t/Q .) in R */
        orc_lc_apply(c->r2q, sc + (size_t)p * L * n, tmpR, NULL, n);               /* 6. exact R -> Q      */
    }
    free(ext); free(d); free(tmpR);
}

/* multiply without relinearisation: out3 = [3][L][n] */
void orc_bfv_multiply(const orc_bfv *c, const u64 *cta, const u64 *ctb, u64 *out3) { tensor_scale(c, cta, ctb, out3); }

/* relinearize a 3-component ciphertext [3][L][n] -> [2][L][n]: hybrid key switching of c2, added onto (c0, c1) */
void orc_bfv_relinearize(const orc_bfv *c, const u64 *ct3, const u64 *rlk, u64 *out) {
    size_t ln = (size_t)c->L * c->n;
    key_switch(c, ct3 + 2 * ln, rlk, ct3, ct3 + ln, out);
}

/* fused form; stage outputs are optional (NULL to skip) so tests can pin each step:
 *   d_scaled: [3][L][n]  the 3-component ciphertext before relinearisation
 *   out     : [2][L][n]  relinearised ciphertext, coefficient form */
void orc_bfv_multiply_relin(const orc_bfv *c, const u64 *cta, const u64 *ctb, const u64 *rlk, u64 *out, u64 *d_scaled) {
    size_t ln = (size_t)c->L * c->n;
    u64 *sc = (u64 *)malloc(3 * ln * sizeof(u64));
    tensor_scale(c, cta, ctb, sc);
    if (d_scaled) memcpy(d_scaled, sc, 3 * ln * sizeof(u64));
    orc_bfv_relinearize(c, sc, rlk, out);
    free(sc);
}

/* ---------- plaintext operands and subtraction (declared only in the reference: include/fhe.cuh:98-104) -------------- */
void orc_bfv_sub(const orc_bfv *c, const u64 *a, const u64 *b, u64 *out) {
    u32 n = c->n, L = c->L;
    for (u32 p = 0; p < 2 * L; p++) { u64 q = c->primes[p % L];
        for (u32 j = 0; j < n; j++) out[(size_t)p * n + j] = orc_submod(a[(size_t)p * n + j], b[(size_t)p * n + j], q); }
}
/* c0 +- Delta*m, c1 unchanged */
void orc_bfv_add_plain(const orc_bfv *c, const u64 *ct, const u64 *pt, u64 *out, int subtract) {
    u32 n = c->n, L = c->L;
    memcpy(out, ct, 2 * (size_t)L * n * sizeof(u64));
    for (u32 i = 0; i < L; i++) { u64 q = c->primes[i];
        for (u32 j = 0; j < n; j++) {
            u64 dm = orc_mulmod(c->delta[i], pt[j] % q, q);
            u64 *o = out + (size_t)i * n + j;
            *o = subtract ? orc_submod(*o, dm, q) : orc_addmod(*o, dm, q);
        } }
}
/* (c0*m, c1*m) in R_Q, m lifted from [0,t) */
void orc_bfv_multiply_plain(const orc_bfv *c, const u64 *ct, const u64 *pt, u64 *out) {
    u32 n = c->n, L = c->L;
    u64 *m = (u64 *)malloc(n * sizeof(u64));
    for (u32 i = 0; i < L; i++) {
        u64 q = c->primes[i];
        for (u32 j = 0; j < n; j++) m[j] = pt[j] % q;
        fwd_limb(c, m, i);
        for (int p = 0; p < 2; p++) {
            u64 *o = out + ((size_t)p * L + i) * n;
            memcpy(o, ct + ((size_t)p * L + i) * n, n * sizeof(u64));
            fwd_limb(c, o, i);
            for (u32 j = 0; j < n; j++) o[j] = orc_mulmod(o[j], m[j], q);
            inv_limb(c, o, i);
        }
    }
    free(m);
}

/* ---------- Galois automorphisms and rotations (declared only in the reference: include/fhe.cuh:59-61,86,113-116) -------- */
/* out(x) = in(x^g) in Z_q[x]/(x^n+1), g odd: coefficient i moves to i*g mod 2n, negated when that index is >= n */
void orc_apply_galois_poly(const u64 *in, u64 *out, u32 n, u32 g, u64 q) {
    for (u32 i = 0; i < n; i++) {
        u32 t = (u32)(((u64)i * g) & (2 * (u64)n - 1));
        if (t < n) out[t] = in[i]; else out[t - n] = in[i] ? q - in[i] : 0;
    }
}
/* gk: [dnum][2][L+K][n] NTT form;  b_d = -a_d s + e_d + P*B_d*s(x^g): switches a ciphertext component under s(x^g) back to s.
 * Streams as for the relinearisation key (so seeds must differ between keys). */
void orc_bfv_galois_keygen(const orc_bfv *c, u64 seed, u32 g, const u64 *sk_ntt, const u64 *cdt, u32 cdt_len, u64 *gk) {
    u32 n = c->n, L = c->L, K = c->K, W = L + K;
    int64_t *e = (int64_t *)malloc(n * sizeof(int64_t));
    u64 *sg = (u64 *)malloc((size_t)W * n * sizeof(u64)), *tmp = (u64 *)malloc(n * sizeof(u64));
    for (u32 i = 0; i < W; i++) {                                  /* s(x^g) per limb, NTT form */
        memcpy(tmp, sk_ntt + (size_t)i * n, n * sizeof(u64));
        inv_limb(c, tmp, i);
        orc_apply_galois_poly(tmp, sg + (size_t)i * n, n, g, c->primes[i]);
        fwd_limb(c, sg + (size_t)i * n, i);
    }
    for (u32 d = 0; d < c->dnum; d++) {
        orc_sample_gaussian(e, n, seed, ST_RLK_BASE * (d + 1) + ST_RLK_E, cdt, cdt_len);
        for (u32 i = 0; i < W; i++) {
            u64 q = c->primes[i];
            u64 *b = gk + ((size_t)(d * 2 + 0) * W + i) * n, *a = gk + ((size_t)(d * 2 + 1) * W + i) * n;
            const u64 *s = sk_ntt + (size_t)i * n, *t = sg + (size_t)i * n;
            orc_sample_uniform(a, n, q, seed, ST_RLK_BASE * (d + 1) + i);
            small_to_limb(b, e, n, q); fwd_limb(c, b, i);
            int in_group = (i < L) && (i / c->alpha == d);
            u64 f = in_group ? c->p_mod_q[i] : 0;
            for (u32 j = 0; j < n; j++) {
                u64 v = orc_submod(b[j], orc_mulmod(a[j], s[j], q), q);
                if (f) v = orc_addmod(v, orc_mulmod(f, t[j], q), q);
                b[j] = v;
            }
        }
    }
    free(e); free(sg); free(tmp);
}
/* ct(x) -> ct(x^g) re-encrypted under s: (c0(x^g), 0) + KeySwitch(c1(x^g)).  ct, out: [2][L][n] coefficient form */
void orc_bfv_apply_galois(const orc_bfv *c, const u64 *ct, u32 g, const u64 *gk, u64 *out) {
    u32 n = c->n, L = c->L;
    u64 *r = (u64 *)malloc(2 * (size_t)L * n * sizeof(u64));
    for (u32 p = 0; p < 2 * L; p++) orc_apply_galois_poly(ct + (size_t)p * n, r + (size_t)p * n, n, g, c->primes[p % L]);
    key_switch(c, r + (size_t)L * n, gk, r, NULL, out);
    free(r);
}
/* mod_switch_to_next (include/fhe.cuh:109): both components lose the last limb of Q with rounding; the result is a ciphertext
 * of the same plaintext for the context built on the first L-1 limbs.  ct: [2][L][n] -> out: [2][L-1][n] */
void orc_modswitch_drop_last(u64 *out, const u64 *in, u32 n, const u64 *moduli, u32 limbs);
/* mod_switch_to_level (include/fhe.cuh:110, declared only): drop the last `drop` limbs of Q one at a time, rounding at every step;
 * out: [2][L-drop][n].  drop = 0 copies. */
void orc_bfv_mod_switch_to_level(const orc_bfv *c, const u64 *ct, u32 drop, u64 *out) {
    u32 n = c->n, L = c->L;
    u64 *cur = (u64 *)malloc((size_t)L * n * sizeof(u64)), *nxt = (u64 *)malloc((size_t)L * n * sizeof(u64));
    for (int p = 0; p < 2; p++) {
        memcpy(cur, ct + (size_t)p * L * n, (size_t)L * n * sizeof(u64));
        for (u32 k = 0; k < drop; k++) { orc_modswitch_drop_last(nxt, cur, n, c->primes, L - k); u64 *t = cur; cur = nxt; nxt = t; }
        memcpy(out + (size_t)p * (L - drop) * n, cur, (size_t)(L - drop) * n * sizeof(u64));
    }
    free(cur); free(nxt);
}

void orc_bfv_mod_switch_to_next(const orc_bfv *c, const u64 *ct, u64 *out) {
    u32 n = c->n, L = c->L;
    for (int p = 0; p < 2; p++) orc_modswitch_drop_last(out + (size_t)p * (L - 1) * n, ct + (size_t)p * L * n, n, c->primes, L);
}
