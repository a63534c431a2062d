#include "rns_consts.hpp"
#include "host_math.hpp"

namespace fhe_b200 {
using namespace host;

// product of basis[l], l != skip, reduced modulo m (skip < 0: whole basis)
static u64 basis_product_mod(const uint64_t* basis, uint32_t count, int skip, u64 m) {
    u64 acc = 1 % m;
    for (uint32_t l = 0; l < count; l++) {
        if ((int)l == skip) continue;
        acc = mulmod(acc, basis[l] % m, m);
    }
    return acc;
}

static void size_consts(LincombConsts& c, const uint64_t* src, uint32_t S, const uint64_t* dst, uint32_t T) {
    c.S = S; c.T = T;
    c.src_mod.assign(src, src + S); c.dst_mod.assign(dst, dst + T);
    c.pre.assign(S, 0); c.th_hi.assign(S, 0); c.th_lo.assign(S, 0);
    c.M.assign((size_t)S * T, 0); c.c.assign(T, 0); c.lam.assign(T, 0);
}

LincombConsts make_conv_consts(const uint64_t* src, uint32_t S, const uint64_t* dst, uint32_t T) {
    LincombConsts c;
    size_consts(c, src, S, dst, T);
    c.use_pre = true;
    for (uint32_t i = 0; i < S; i++) {
        const u64 qi = src[i];
        c.pre[i] = invmod(basis_product_mod(src, S, (int)i, qi), qi);
        frac128(1, qi, c.th_hi[i], c.th_lo[i]);                       // theta_i = 1/q_i
        for (uint32_t k = 0; k < T; k++) c.M[(size_t)i * T + k] = basis_product_mod(src, S, (int)i, dst[k]);
    }
    for (uint32_t k = 0; k < T; k++) {
        const u64 Qm = basis_product_mod(src, S, -1, dst[k]);
        c.c[k] = Qm ? dst[k] - Qm : 0;                                 // -Q mod m_k
    }
    return c;
}

LincombConsts make_scale_consts(const uint64_t* qs, uint32_t L, const uint64_t* ps, uint32_t R, uint64_t t,
                                const uint64_t* targets, uint32_t T, bool with_extra) {
    LincombConsts c;
    size_consts(c, qs, L, targets, T);
    c.use_extra = with_extra;
    for (uint32_t i = 0; i < L; i++) {
        const u64 qi = qs[i];
        const u64 P_mod_qi = basis_product_mod(ps, R, -1, qi);
        // Qt = ((Q P)/q_i)^-1 mod q_i ;  t Qt P / q_i = omega_i + rho_i / q_i
        const u64 Qt = invmod(mulmod(basis_product_mod(qs, L, (int)i, qi), P_mod_qi, qi), qi);
        const u64 rho = mulmod(mulmod(t % qi, Qt, qi), P_mod_qi, qi);
        frac128(rho, qi, c.th_hi[i], c.th_lo[i]);
        for (uint32_t k = 0; k < T; k++) {
            const u64 m = targets[k];
            const u64 A = mulmod(mulmod(t % m, Qt % m, m), basis_product_mod(ps, R, -1, m), m);   // t Qt P mod m
            c.M[(size_t)i * T + k] = mulmod(submod(A, rho % m, m), invmod(qi % m, m), m);          // omega_i mod m
        }
    }
    for (uint32_t k = 0; k < T; k++) {
        const u64 m = targets[k];
        c.c[k] = 1 % m;
        if (with_extra) {
            int j = -1;                                   // position of this target inside the P basis
            for (uint32_t r = 0; r < R; r++) if (ps[r] == m) j = (int)r;
            if (j < 0) { c.S = 0; return c; }             // caller reports the error
            const u64 over = mulmod(basis_product_mod(qs, L, -1, m), basis_product_mod(ps, R, j, m), m);
            c.lam[k] = mulmod(mulmod(t % m, invmod(over, m), m), basis_product_mod(ps, R, j, m), m);
        }
    }
    return c;
}

}  // namespace fhe_b200
