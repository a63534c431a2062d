#include "tables.hpp"

namespace fhe_b200 {

int build_limb_tables(uint64_t q, uint32_t n, Twiddle* fwd, Twiddle* inv, LimbParams* P) {
    using namespace host;
    if (q < 3 || (q >> 61) != 0 || (q & 1) == 0 || !is_prime(q)) return -1;
    const u64 psi = find_psi(q, n);
    if (psi == 0) return -1;
    const u32 lg = ilog2(n);
    const u64 ipsi = invmod(psi, q);
    u64 p = 1, ip = 1;
    for (u32 e = 0; e < n; e++) {
        const u32 k = bitrev(e, lg);
        fwd[k].w = p; fwd[k].ws = shoup(p, q);
        inv[k].w = ip; inv[k].ws = shoup(ip, q);
        p = mulmod(p, psi, q); ip = mulmod(ip, ipsi, q);
    }
    P->q = q;
    frac128(1, q, P->mu_hi, P->mu_lo);
    P->ninv = invmod(n % q, q);
    P->ninv_s = shoup(P->ninv, q);
    P->w1ninv = mulmod(inv[1].w, P->ninv, q);
    P->w1ninv_s = shoup(P->w1ninv, q);
    u64 bits = 0; for (u64 t = q; t; t >>= 1) bits++;
    P->qbits = bits;
    return 0;
}

int lazy_headroom(const uint64_t* moduli, uint32_t count) {
    int hb = 16;
    for (uint32_t i = 0; i < count; i++) if (moduli[i] >> 60) hb = 8;
    return hb;
}

}  // namespace fhe_b200
