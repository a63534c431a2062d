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
    P->nm = ((q - 1) >> lg) | ((u64)lg << 56);
    P->w1ninv = mulmod(inv[1].w, P->ninv, q);
    P->w1ninv_s = shoup(P->w1ninv, q);
    P->tq = (u64)kTQ * q;
    return 0;
}

size_t tile_p3_entries(uint32_t logn) {
    const int LB = tile_lb(logn);
    return (size_t)(1u << (LB - 4)) * (16 - (1u << (12 - LB)));
}

void build_tile_tables(const Twiddle* main, uint32_t logn, Twiddle* p12, Twiddle* p3) {
    const int LB = tile_lb(logn), K1 = (int)logn - LB, R3 = LB - 8;
    const uint32_t NT = 1u << (LB - 4), KB = 1u << (4 - R3);
    const size_t p3n = tile_p3_entries(logn);
    for (uint32_t b = 0; b < (1u << K1); b++) {
        const uint32_t root = (1u << K1) + b;
        Twiddle* a = p12 + (size_t)b * 256;
        a[255].w = 0; a[255].ws = 0;
        for (int v = 0; v < 4; v++)
            for (uint32_t key = 0; key < (1u << v); key++) a[(1u << v) - 1 + key] = main[(root << v) + key];
        for (int v = 0; v < 4; v++)
            for (uint32_t hi = 0; hi < 16; hi++)
                for (uint32_t key = 0; key < (1u << v); key++)
                    a[15 + 16 * ((1u << v) - 1) + (hi << v) + key] = main[((((root << 4) + hi)) << v) + key];
        Twiddle* c = p3 + (size_t)b * p3n;
        for (int v = 0; v < R3; v++)
            for (uint32_t key = 0; key < (KB << v); key++)
                for (uint32_t tid = 0; tid < NT; tid++)
                    c[(size_t)NT * KB * ((1u << v) - 1) + (size_t)key * NT + tid] =
                        main[((((root << (4 + R3)) + tid)) << (4 - R3 + v)) + key];
    }
}

void build_bal_tables(const Twiddle* main, uint32_t logn, Twiddle* out) {
    const uint32_t ka = logn - 8, pairs = 1u << (ka - 1);
    for (uint32_t p = 0; p < pairs; p++) {
        Twiddle* a = out + (size_t)p * 512;
        for (uint32_t t = 0; t < 2; t++) {
            const uint32_t root = (1u << ka) + 2 * p + t;
            a[t * 16 + 15].w = 0; a[t * 16 + 15].ws = 0;
            for (int v = 0; v < 4; v++)
                for (uint32_t key = 0; key < (1u << v); key++) a[t * 16 + (1u << v) - 1 + key] = main[(root << v) + key];
            for (int v = 0; v < 4; v++)
                for (uint32_t key = 0; key < (1u << v); key++)
                    for (uint32_t e = 0; e < 16; e++)
                        a[32 + 32 * ((1u << v) - 1) + key * 32 + t * 16 + e] = main[(((root << 4) + e) << v) + key];
        }
    }
}

int lazy_headroom(const uint64_t* moduli, uint32_t count) {
    int hb = 16;
    for (uint32_t i = 0; i < count; i++) if (moduli[i] >> 60) hb = 8;
    return hb;
}

bool all_near60(const uint64_t* moduli, uint32_t count) {
    for (uint32_t i = 0; i < count; i++)
        if ((moduli[i] >> 60) != 0 || moduli[i] <= (1ull << 60) - (1ull << 32)) return false;
    return true;
}

}  // namespace fhe_b200
