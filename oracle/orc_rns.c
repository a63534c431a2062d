/*
 * oracle/orc_rns.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * CPU restatement of the reference's RNS layer *as documented* (its code is
 * placeholder: to_rns copies, from_rns writes 0, fast_base_conversion_kernel
 * and rns_mod_switch_kernel are declared only):
 *   CRT / RNS formulas            /root/reference/docs/ARCHITECTURE.md:161-187, include/rns.cuh:10-19
 *   fast base conversion (Bajard) /root/reference/include/rns.cuh:116-125
 *       out[k] = sum_i [x_i * (Q/q_i)^-1]_{q_i} * (Q/q_i mod m_k)
 *   modulus switching             /root/reference/include/rns.cuh:44-45,128-136, docs/ARCHITECTURE.md:245-248
 *
 * All conversions here are EXACT: the overflow count v = round(sum_i y_i/q_i)
 * is computed in 128-bit integer fixed point (no floating point), so
 * the result equals the centred representative of x reduced mod each target,
 * except when x/Q is within 2^-57 of +-1/2.  The same integer recipe is what
 * the CUDA engine implements, so parity is bit-exact; tests additionally
 * check the recipe against Python big integers.
 *
 * One primitive covers every use (Q->R extension, t/Q scale-and-round,
 * R->Q, key-switch ModUp / ModDown, decryption scaling):
 *
 *   z_i   = use_pre ? x_i * pre_i mod s_i : x_i
 *   I     = ( sum_i ( z_i * Th_i  +  hi64(z_i * Tl_i) )  +  2^63 ) >> 64        (192-bit sum)
 *   out_k = ( sum_i z_i * M[i][k]  +  (I mod m_k) * c_k  +  extra_k * lam_k ) mod m_k
 */
#include "orc_math.h"
#include <stdlib.h>
#include <string.h>

typedef struct {
    u32 S, T;
    u64 *src_mod, *dst_mod;
    int use_pre;  u64 *pre;
    u64 *th_hi, *th_lo;
    u64 *M;            /* [S][T] */
    u64 *c;            /* [T] */
    int use_extra; u64 *lam; /* [T] */
} orc_lincomb;

/* extended Euclid inverse, works for any coprime pair (t may be a power of two) */
static u64 inv_general(u64 a, u64 m) {
    __int128 r0 = m, r1 = a % m, t0 = 0, t1 = 1;
    while (r1) { __int128 q = r0 / r1, r2 = r0 - q * r1, t2 = t0 - q * t1; r0 = r1; r1 = r2; t0 = t1; t1 = t2; }
    if (r0 != 1) return 0;
    if (t0 < 0) t0 += m;
    return (u64)t0;
}
u64 orc_inv_general(u64 a, u64 m) { return inv_general(a, m); }

static orc_lincomb *lc_alloc(u32 S, u32 T, const u64 *src, const u64 *dst) {
    orc_lincomb *lc = (orc_lincomb *)calloc(1, sizeof(*lc));
    lc->S = S; lc->T = T;
    lc->src_mod = (u64 *)malloc(S * sizeof(u64)); memcpy(lc->src_mod, src, S * sizeof(u64));
    lc->dst_mod = (u64 *)malloc(T * sizeof(u64)); memcpy(lc->dst_mod, dst, T * sizeof(u64));
    lc->pre = (u64 *)calloc(S, sizeof(u64));
    lc->th_hi = (u64 *)calloc(S, sizeof(u64)); lc->th_lo = (u64 *)calloc(S, sizeof(u64));
    lc->M = (u64 *)calloc((size_t)S * T, sizeof(u64));
    lc->c = (u64 *)calloc(T, sizeof(u64)); lc->lam = (u64 *)calloc(T, sizeof(u64));
    return lc;
}
void orc_lc_free(orc_lincomb *lc) {
    if (!lc) return;
    free(lc->src_mod); free(lc->dst_mod); free(lc->pre); free(lc->th_hi); free(lc->th_lo);
    free(lc->M); free(lc->c); free(lc->lam); free(lc);
}

/* prod_{l != skip} src[l] mod m   (skip = -1: full product) */
static u64 prod_mod(const u64 *src, u32 S, int skip, u64 m) {
    u64 r = 1 % m;
    for (u32 l = 0; l < S; l++) if ((int)l != skip) r = orc_mulmod(r, src[l] % m, m);
    return r;
}

/* exact centred base conversion  src basis (product Q) -> dst moduli  (rns.cuh:116-125 intent) */
orc_lincomb *orc_lc_make_conv(const u64 *src, u32 S, const u64 *dst, u32 T) {
    orc_lincomb *lc = lc_alloc(S, T, src, dst);
    lc->use_pre = 1;
    for (u32 i = 0; i < S; i++) {
        u64 qi = src[i];
        lc->pre[i] = inv_general(prod_mod(src, S, (int)i, qi), qi);   /* (Q/q_i)^-1 mod q_i */
        orc_frac128(1, qi, &lc->th_hi[i], &lc->th_lo[i]);            /* 2^128 / q_i       */
        for (u32 k = 0; k < T; k++) lc->M[(size_t)i * T + k] = prod_mod(src, S, (int)i, dst[k]);
    }
    for (u32 k = 0; k < T; k++) lc->c[k] = orc_negmod(prod_mod(src, S, -1, dst[k]), dst[k]);  /* -Q mod m_k */
    return lc;
}

/* out = round(t/Q * d) over targets.
 * qs[L]: basis the division is by; ps[R]: complementary basis d is also known in (R may be 0);
 * targets: either ps itself (with_extra=1: HPS BFV scale step) or arbitrary moduli coprime to Q
 * with R == 0 (decryption: targets = {t}). */
orc_lincomb *orc_lc_make_scale(const u64 *qs, u32 L, const u64 *ps, u32 R, u64 t,
                               const u64 *targets, u32 T, int with_extra) {
    orc_lincomb *lc = lc_alloc(L, T, qs, targets);
    lc->use_pre = 0;
    lc->use_extra = with_extra;
    for (u32 i = 0; i < L; i++) {
        u64 qi = qs[i];
        /* Qt_i = ((Q*P)/q_i)^-1 mod q_i ;  rho_i = t * Qt_i * P mod q_i */
        u64 qp_over = orc_mulmod(prod_mod(qs, L, (int)i, qi), prod_mod(ps, R, -1, qi), qi);
        u64 Qt = inv_general(qp_over, qi);
        u64 rho = orc_mulmod(orc_mulmod(t % qi, Qt, qi), prod_mod(ps, R, -1, qi), qi);
        orc_frac128(rho, qi, &lc->th_hi[i], &lc->th_lo[i]);
        for (u32 k = 0; k < T; k++) {
            u64 m = targets[k];
            /* omega_{i,k} = (t*Qt*P - rho)/q_i mod m */
            u64 A = orc_mulmod(orc_mulmod(t % m, Qt % m, m), prod_mod(ps, R, -1, m), m);
            u64 num = orc_submod(A, rho % m, m);
            lc->M[(size_t)i * T + k] = orc_mulmod(num, inv_general(qi % m, m), m);
        }
    }
    for (u32 k = 0; k < T; k++) {
        u64 m = targets[k];
        lc->c[k] = 1 % m;
        if (with_extra) {
            /* lambda_k = t * ((Q*P)/p_j)^-1 * (P/p_j) mod p_j, j = position of the target inside ps (targets subset of ps) */
            int j = -1;
            for (u32 r = 0; r < R; r++) if (ps[r] == m) j = (int)r;
            u64 qp_over = orc_mulmod(prod_mod(qs, L, -1, m), prod_mod(ps, R, j, m), m);
            lc->lam[k] = orc_mulmod(orc_mulmod(t % m, inv_general(qp_over, m), m), prod_mod(ps, R, j, m), m);
        }
    }
    return lc;
}

/* in: [S][n], extra: [T][n] or NULL, out: [T][n] */
void orc_lc_apply(const orc_lincomb *lc, u64 *out, const u64 *in, const u64 *extra, u32 n) {
    u32 S = lc->S, T = lc->T;
    #pragma omp parallel for schedule(static)
    for (u32 j = 0; j < n; j++) {
        u64 z[128];
        u128 acc = 0; u64 carry = 0;
        for (u32 i = 0; i < S; i++) {
            u64 x = in[(size_t)i * n + j];
            z[i] = lc->use_pre ? orc_mulmod(x, lc->pre[i], lc->src_mod[i]) : x;
            u128 term = (u128)z[i] * lc->th_hi[i] + (u64)(((u128)z[i] * lc->th_lo[i]) >> 64);
            u128 s = acc + term; if (s < acc) carry++; acc = s;
        }
        { u128 s = acc + ((u128)1 << 63); if (s < acc) carry++; acc = s; }
        u128 I = ((u128)carry << 64) | (u64)(acc >> 64);
        for (u32 k = 0; k < T; k++) {
            u64 m = lc->dst_mod[k];
            u128 sum = 0;                       /* every term < 2^122, S+2 <= 64 terms: no overflow */
            for (u32 i = 0; i < S; i++) sum += (u128)z[i] * lc->M[(size_t)i * T + k];
            sum += (u128)(u64)(I % m) * lc->c[k];
            if (lc->use_extra) sum += (u128)extra[(size_t)k * n + j] * lc->lam[k];
            out[(size_t)k * n + j] = (u64)(sum % m);
        }
    }
}

/* accessors so tests can compare constants with big-integer recomputation and with the CUDA engine */
u32 orc_lc_S(const orc_lincomb *lc) { return lc->S; }
u32 orc_lc_T(const orc_lincomb *lc) { return lc->T; }
void orc_lc_get(const orc_lincomb *lc, u64 *pre, u64 *th_hi, u64 *th_lo, u64 *M, u64 *c, u64 *lam) {
    memcpy(pre, lc->pre, lc->S * sizeof(u64));
    memcpy(th_hi, lc->th_hi, lc->S * sizeof(u64)); memcpy(th_lo, lc->th_lo, lc->S * sizeof(u64));
    memcpy(M, lc->M, (size_t)lc->S * lc->T * sizeof(u64));
    memcpy(c, lc->c, lc->T * sizeof(u64)); memcpy(lam, lc->lam, lc->T * sizeof(u64));
}

/* RNSContext::to_rns (rns.cuh:33): residues of 64-bit values, limb-major [k][count] */
void orc_to_rns(u64 *out, const u64 *values, u32 count, const u64 *moduli, u32 k) {
    for (u32 l = 0; l < k; l++) for (u32 j = 0; j < count; j++) out[(size_t)l * count + j] = values[j] % moduli[l];
}

/* RNS modulus switch, drop the last limb with rounding (rns.cuh:128-136 intent):
 *   out_i = (x_i - [x_last]_centred) * q_last^-1 mod q_i  ==  round(x / q_last) mod q_i */
void orc_modswitch_drop_last(u64 *out, const u64 *in, u32 n, const u64 *moduli, u32 limbs) {
    u64 ql = moduli[limbs - 1], half = ql >> 1;
    for (u32 i = 0; i + 1 < limbs; i++) {
        u64 qi = moduli[i], inv = inv_general(ql % qi, qi);
        for (u32 j = 0; j < n; j++) {
            u64 r = in[(size_t)(limbs - 1) * n + j];
            /* centred remainder: r in (-ql/2, ql/2] -> subtract r, or r-ql */
            u64 rr = r > half ? orc_submod(r % qi, ql % qi, qi) : r % qi;
            out[(size_t)i * n + j] = orc_mulmod(orc_submod(in[(size_t)i * n + j], rr, qi), inv, qi);
        }
    }
}
