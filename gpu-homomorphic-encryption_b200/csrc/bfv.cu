// BFV on RNS: the path under fhe::FHEContext (keygen, relinkey_gen, encrypt, decrypt, add, multiply+relinearize).
//
// Scheme equations follow /root/reference/src/fhe.cu:54-235 and docs/ARCHITECTURE.md:260-326; the parts the
// reference leaves out (honouring log_q through an RNS basis, the t/Q scaling of the tensor product, and
// relinearisation -- `relinearize` there is components.resize(2)) are built as
//   HPS-style RNS multiplication: exact extension Q -> Q u R, tensor in the NTT domain, round(t/Q .) computed in R,
//   exact conversion back R -> Q,  then hybrid (dnum-digit) key switching with special modulus P = first K primes of R:
//   ModUp (exact) -> NTT -> inner product with the key -> INTT -> ModDown (exact, fused with the final add).
// All base conversions go through the lincomb kernel; every NTT through launch_ntt.
//
// Buffers (uint64):  sk [L+R][N] NTT | pk [2][L][N] NTT | rlk [dnum][2][L+K][N] NTT | ct [B][2][L][N] coeff | pt [B][N].
#include "common.cuh"
#include "lincomb.cuh"
#include "bfv.cuh"
#include "samplers.cuh"
#include "host_math.hpp"
#include <cmath>
#include <cstring>


namespace fhe_b200 {
int ensure_words(uint64_t** buf, size_t* have, size_t need) {
    if (*have >= need) return 0;
    if (*buf) { FHE_CUDA(cudaDeviceSynchronize()); FHE_CUDA(cudaFree(*buf)); *buf = nullptr; *have = 0; }
    FHE_CUDA(cudaMalloc(buf, need * sizeof(uint64_t)));
    *have = need;
    return 0;
}
}  // namespace fhe_b200

namespace fhe_b200 {

// ---- sampling kernels ----------------------------------------------------------------------------------------------
// small polynomial (ternary or Gaussian), expanded to residues for limbs [limb_begin, +limb_count).  item0 < 0: one polynomial
// keyed by the seed itself (keys); item0 >= 0: polynomial b is batch item item0 + b of the call and uses rng_item_seed(seed, item0 + b)
template <int MODE>   // 0 ternary, 1 gaussian
__global__ void __launch_bounds__(256) sample_small_kernel(u64* __restrict__ out, const LimbParams* __restrict__ params,
                                                           uint32_t logn, uint32_t limb_begin, uint32_t limb_count, uint32_t batch,
                                                           u64 seed, long long item0, u64 stream, u32 thr, const u64* __restrict__ cdt, u32 cdt_len,
                                                           const RngKey rk) {
    __shared__ u64 scdt[kMaxCdt];
    if (MODE == 1) { for (u32 i = threadIdx.x; i < cdt_len; i += blockDim.x) scdt[i] = cdt[i]; __syncthreads(); }
    const uint32_t n = 1u << logn;
    const size_t total = (size_t)batch * n;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(g & (n - 1));
        const size_t b = g >> logn;
        const u64 r = rng_word(rk, item0 < 0 ? seed : rng_item_seed(seed, (u64)item0 + b), stream, j);
        const int v = MODE == 0 ? ternary_from(r, thr) : gauss_from(r, scdt, cdt_len);
        for (uint32_t l = 0; l < limb_count; l++) out[(b * limb_count + l) * n + j] = small_to_residue(v, params[limb_begin + l].q);
    }
}
// host-generated small polynomial (int8) -> residues
__global__ void __launch_bounds__(256) expand_small_kernel(u64* __restrict__ out, const int8_t* __restrict__ s,
                                                           const LimbParams* __restrict__ params, uint32_t n, uint32_t limb_begin, uint32_t limb_count) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int v = s[j];
        for (uint32_t l = 0; l < limb_count; l++) out[(size_t)l * n + j] = small_to_residue(v, params[limb_begin + l].q);
    }
}
// uniform residues: limb l uses stream stream_base + l
__global__ void __launch_bounds__(256) sample_uniform_kernel(u64* __restrict__ out, const LimbParams* __restrict__ params, uint32_t logn,
                                                             uint32_t limb_begin, uint32_t limb_count, u64 seed, u64 stream_base, const RngKey rk) {
    const uint32_t n = 1u << logn;
    const size_t total = (size_t)limb_count * n;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(g & (n - 1));
        const uint32_t l = (uint32_t)(g >> logn);
        const LimbParams P = params[limb_begin + l];
        out[g] = barrett128(rng_word(rk, seed, stream_base + l, 2ull * j), rng_word(rk, seed, stream_base + l, 2ull * j + 1), P.q, P.mu_hi, P.mu_lo);
    }
}

// ---- key generation -------------------------------------------------------------------------------------------------
// pk0 = e - a*s  (all NTT form), one polynomial of L limbs
__global__ void __launch_bounds__(256) pk_finish_kernel(u64* __restrict__ pk0, const u64* __restrict__ a, const u64* __restrict__ s,
                                                        const LimbParams* __restrict__ params, uint32_t logn, size_t total) {
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const LimbParams P = params[g >> logn];
        pk0[g] = sub_mod(pk0[g], mul_mod(a[g], s[g], P), P.q);
    }
}
// b = e - a*s + f*s^2, f = P mod q_i for the digit's own limbs, 0 elsewhere; W = L+K limbs
// (s_from: the key being switched FROM in NTT form -- s(x^g) for a Galois key; nullptr = s^2, the relinearisation key)
__global__ void __launch_bounds__(256) rlk_finish_kernel(u64* __restrict__ b, const u64* __restrict__ a, const u64* __restrict__ s,
                                                         const u64* __restrict__ s_from,
                                                         const LimbParams* __restrict__ params, const u64* __restrict__ pmodq,
                                                         uint32_t logn, uint32_t g_lo, uint32_t g_hi, size_t total) {
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(g >> logn);
        const LimbParams P = params[i];
        const u64 sv = s[g];
        u64 v = sub_mod(b[g], mul_mod(a[g], sv, P), P.q);
        if (i >= g_lo && i < g_hi) v = add_mod(v, mul_mod(pmodq[i], s_from ? s_from[g] : mul_mod(sv, sv, P), P), P.q);
        b[g] = v;
    }
}

// out(x) = in(x^g) per limb (g odd, ginv = g^-1 mod 2N): output coefficient j comes from i' = j*ginv mod 2N, negated when i' >= N.
// Buffers [polys][limb_count][N], limb l of the buffer = plan limb limb_begin + l.  Gather form: coalesced stores.
__global__ void __launch_bounds__(256) galois_kernel(u64* __restrict__ out, const u64* __restrict__ in,
                                                     const LimbParams* __restrict__ params, uint32_t logn, uint32_t limb_begin,
                                                     uint32_t limb_count, uint32_t ginv, size_t total) {
    const uint32_t n = 1u << logn;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(g & (n - 1));
        const size_t pl = g >> logn;
        const u64 q = params[limb_begin + (uint32_t)(pl % limb_count)].q;
        const uint32_t ip = (uint32_t)(((u64)j * ginv) & (2 * (u64)n - 1));
        const u64 v = in[(pl << logn) + (ip & (n - 1))];
        out[g] = ip < n ? v : neg_mod(v, q);
    }
}

// ---- encrypt ----------------------------------------------------------------------------------------------------------
// ct[b][0][i] = pk0[i]*u[b][i], ct[b][1][i] = pk1[i]*u[b][i]   (NTT form)
__global__ void __launch_bounds__(256) enc_mul_kernel(u64* __restrict__ ct, const u64* __restrict__ u, const u64* __restrict__ pk,
                                                      const LimbParams* __restrict__ params, uint32_t logn, uint32_t L, size_t total) {
    const size_t ln = (size_t)L << logn;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const size_t b = g / ln, r = g % ln;
        const LimbParams P = params[r >> logn];
        const u64 uv = u[g];
        ct[(2 * b) * ln + r] = mul_mod(pk[r], uv, P);
        ct[(2 * b + 1) * ln + r] = mul_mod(pk[ln + r], uv, P);
    }
}
// c0 += e1 + delta*m ; c1 += e2   (coefficient form; e1, e2 regenerated from the counter generator)
__global__ void __launch_bounds__(256) enc_finish_kernel(u64* __restrict__ ct, const u64* __restrict__ pt, const LimbParams* __restrict__ params,
                                                         const u64* __restrict__ delta, const u64* __restrict__ cdt, u32 cdt_len,
                                                         uint32_t logn, uint32_t L, uint32_t batch, u64 seed, u64 item0, const RngKey rk) {
    __shared__ u64 scdt[kMaxCdt];
    for (u32 i = threadIdx.x; i < cdt_len; i += blockDim.x) scdt[i] = cdt[i];
    __syncthreads();
    const uint32_t n = 1u << logn;
    const size_t total = (size_t)batch * n;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(g & (n - 1));
        const size_t b = g >> logn;
        const u64 sd = rng_item_seed(seed, item0 + b);
        const int e1 = gauss_from(rng_word(rk, sd, 1, j), scdt, cdt_len);
        const int e2 = gauss_from(rng_word(rk, sd, 2, j), scdt, cdt_len);
        const u64 m = pt[g];
        for (uint32_t i = 0; i < L; i++) {
            const LimbParams P = params[i];
            const size_t o0 = ((2 * b) * L + i) * n + j, o1 = ((2 * b + 1) * L + i) * n + j;
            const u64 dm = mul_mod(delta[i], barrett128(0, m, P.q, P.mu_hi, P.mu_lo), P);
            ct[o0] = add_mod(add_mod(ct[o0], small_to_residue(e1, P.q), P.q), dm, P.q);
            ct[o1] = add_mod(ct[o1], small_to_residue(e2, P.q), P.q);
        }
    }
}

// ---- plaintext operands ------------------------------------------------------------------------------------------------
// out[b][0][i] = ct[b][0][i] +- delta_i * m[b],  out[b][1][i] = ct[b][1][i]
__global__ void __launch_bounds__(256) add_plain_kernel(u64* __restrict__ out, const u64* __restrict__ ct, const u64* __restrict__ pt,
                                                        const LimbParams* __restrict__ params, const u64* __restrict__ delta,
                                                        uint32_t logn, uint32_t L, size_t total /* B*2*L*N */, int subtract) {
    const uint32_t n = 1u << logn;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(g & (n - 1));
        const size_t pl = g >> logn;                  // (b*2 + comp)*L + i
        const uint32_t i = (uint32_t)(pl % L);
        const size_t bc = pl / L;
        u64 v = ct[g];
        if ((bc & 1) == 0) {
            const LimbParams P = params[i];
            const u64 dm = mul_mod(delta[i], barrett128(0, pt[(bc >> 1) * n + j], P.q, P.mu_hi, P.mu_lo), P);
            v = subtract ? sub_mod(v, dm, P.q) : add_mod(v, dm, P.q);
        }
        out[g] = v;
    }
}
// lifted plaintext: out[b][i][j] = pt[b][j] mod q_i
__global__ void __launch_bounds__(256) lift_plain_kernel(u64* __restrict__ out, const u64* __restrict__ pt, const LimbParams* __restrict__ params,
                                                         uint32_t logn, uint32_t L, size_t total /* B*L*N */) {
    const uint32_t n = 1u << logn;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(g & (n - 1));
        const size_t pl = g >> logn;
        const LimbParams P = params[pl % L];
        out[g] = barrett128(0, pt[(pl / L) * n + j], P.q, P.mu_hi, P.mu_lo);
    }
}
// ct[b][c][i] *= m[b][i]  (NTT form)
__global__ void __launch_bounds__(256) mul_plain_kernel(u64* __restrict__ ct, const u64* __restrict__ m, const LimbParams* __restrict__ params,
                                                        uint32_t logn, uint32_t L, size_t total /* B*2*L*N */) {
    const size_t ln = (size_t)L << logn;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const size_t b = g / (2 * ln), r = g % ln;
        ct[g] = mul_mod(ct[g], m[b * ln + r], params[r >> logn]);
    }
}

// ---- decrypt ----------------------------------------------------------------------------------------------------------
// x[b][i] = NTT(c1)[b][i] * s[i]
__global__ void __launch_bounds__(256) dec_mul_kernel(u64* __restrict__ x, const u64* __restrict__ s, const LimbParams* __restrict__ params,
                                                      uint32_t logn, uint32_t L, size_t total) {
    const size_t ln = (size_t)L << logn;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const size_t r = g % ln;
        x[g] = mul_mod(x[g], s[r], params[r >> logn]);
    }
}
// x[b][i] += c0[b][i]   (ct is [B][2][L][N])
__global__ void __launch_bounds__(256) dec_add_kernel(u64* __restrict__ x, const u64* __restrict__ ct, const LimbParams* __restrict__ params,
                                                      uint32_t logn, uint32_t L, size_t total) {
    const size_t ln = (size_t)L << logn;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const size_t b = g / ln, r = g % ln;
        x[g] = add_mod(x[g], ct[(2 * b) * ln + r], params[r >> logn].q);
    }
}

// ---- multiply ---------------------------------------------------------------------------------------------------------
// ext: [4][B][A][N] (a0,a1,b0,b1; NTT form) -> d: [3][B][A][N]   d0 = a0 b0, d1 = a0 b1 + a1 b0, d2 = a1 b1
__global__ void __launch_bounds__(256) tensor_kernel(ulonglong2* __restrict__ d, const ulonglong2* __restrict__ ext,
                                                     const LimbParams* __restrict__ params, uint32_t logn, uint32_t limb_begin, uint32_t A,
                                                     size_t per_comp /* B*A*N/2 */, size_t b_off /* 2 per_comp; 0 when squaring */) {
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < per_comp; v += (size_t)gridDim.x * blockDim.x) {
        const LimbParams P = params[limb_begin + ((2 * v) >> logn) % A];
        const ulonglong2 a0 = ext[v], a1 = ext[per_comp + v], b0 = ext[b_off + v], b1 = ext[b_off + per_comp + v];
        ulonglong2 r0, r1, r2;
        u64 hi, lo;
        r0.x = mul_mod(a0.x, b0.x, P); r0.y = mul_mod(a0.y, b0.y, P);
        r2.x = mul_mod(a1.x, b1.x, P); r2.y = mul_mod(a1.y, b1.y, P);
        hi = 0; lo = 0; mac128(hi, lo, a0.x, b1.x); mac128(hi, lo, a1.x, b0.x); r1.x = barrett128(hi, lo, P.q, P.mu_hi, P.mu_lo);
        hi = 0; lo = 0; mac128(hi, lo, a0.y, b1.y); mac128(hi, lo, a1.y, b0.y); r1.y = barrett128(hi, lo, P.q, P.mu_hi, P.mu_lo);
        d[v] = r0; d[per_comp + v] = r1; d[2 * per_comp + v] = r2;
    }
}
// key-switch inner product: acc[c][b][i] = sum_dg dig[dg][b][i] * rlk[dg][c][i]    (W = L+K limbs, NTT form)
// A thread owns vector r of the polynomials and a run of `bpt` batch items: the 2 dnum key vectors of r -- six of the eleven 16-byte
// accesses per item when every item fetched them itself -- are loaded once (dnum <= 4: registers) and reused for the whole run.
__global__ void __launch_bounds__(256) ks_inner_kernel(ulonglong2* __restrict__ acc, const ulonglong2* __restrict__ dig,
                                                       const ulonglong2* __restrict__ rlk, const LimbParams* __restrict__ params,
                                                       uint32_t logn, uint32_t limb_begin, uint32_t W, uint32_t dnum, uint32_t batch, uint32_t bpt) {
    const size_t wn = ((size_t)W << logn) / 2;              // vectors per polynomial
    const size_t per = wn * batch;
    const uint32_t runs = (batch + bpt - 1) / bpt;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < wn * runs; v += (size_t)gridDim.x * blockDim.x) {
        const size_t r = v % wn;
        const uint32_t b0 = (uint32_t)(v / wn) * bpt, b1 = min(batch, b0 + bpt);
        const LimbParams P = params[limb_begin + ((2 * r) >> logn)];
        ulonglong2 kb[4], ka[4];
        if (dnum <= 4) {
#pragma unroll
            for (int dg = 0; dg < 4; dg++)
                if ((uint32_t)dg < dnum) { kb[dg] = rlk[(size_t)(2 * dg) * wn + r]; ka[dg] = rlk[(size_t)(2 * dg + 1) * wn + r]; }
        }
        for (uint32_t b = b0; b < b1; b++) {
            const size_t o = (size_t)b * wn + r;
            u64 h0x = 0, l0x = 0, h0y = 0, l0y = 0, h1x = 0, l1x = 0, h1y = 0, l1y = 0;
            if (dnum <= 4) {
#pragma unroll
                for (int dg = 0; dg < 4; dg++)
                    if ((uint32_t)dg < dnum) {
                        const ulonglong2 x = dig[dg * per + o];
                        mac128(h0x, l0x, x.x, kb[dg].x); mac128(h0y, l0y, x.y, kb[dg].y);
                        mac128(h1x, l1x, x.x, ka[dg].x); mac128(h1y, l1y, x.y, ka[dg].y);
                    }
            } else {
                for (uint32_t dg = 0; dg < dnum; dg++) {
                    const ulonglong2 x = dig[dg * per + o];
                    const ulonglong2 k0 = rlk[(size_t)(2 * dg) * wn + r], k1 = rlk[(size_t)(2 * dg + 1) * wn + r];
                    mac128(h0x, l0x, x.x, k0.x); mac128(h0y, l0y, x.y, k0.y);
                    mac128(h1x, l1x, x.x, k1.x); mac128(h1y, l1y, x.y, k1.y);
                }
            }
            ulonglong2 o0, o1;
            o0.x = barrett128(h0x, l0x, P.q, P.mu_hi, P.mu_lo); o0.y = barrett128(h0y, l0y, P.q, P.mu_hi, P.mu_lo);
            o1.x = barrett128(h1x, l1x, P.q, P.mu_hi, P.mu_lo); o1.y = barrett128(h1y, l1y, P.q, P.mu_hi, P.mu_lo);
            acc[o] = o0; acc[per + o] = o1;
        }
    }
}
// batch items per thread: whole runs where one polynomial alone fills the GPU, one item otherwise (small rings live on the batch)
static inline uint32_t ks_inner_bpt(int sm_count, size_t vectors_per_poly, uint32_t batch) {
    if (vectors_per_poly < (size_t)sm_count * 16 * 256 / 2 || batch < 2) return 1;
    return batch < 8 ? batch : 8;
}

// FHE_B200_FUSED_TILE=1: the tensor product and the key-switch inner product run inside the tile passes (ntt_fused.cu; two-pass
// transform, at most three digits).  Bit-identical, six launches fewer per multiply, and MEASURED SLOWER on the B200 (config 4,
// batch 8: 1 303 against 1 474 multiplies/s): every kernel here is bound by the integer pipe, so the HBM round
// trip the fusion removes was already hidden, while three 4 KiB buffers per warp plus both twiddle blocks leave one CTA of 8 warps
// per SM and nothing for the other stream's conversion kernel to co-run with.  Hence opt-in; read at every call.
static bool use_fused_tile(const fhe_b200_bfv* c) {
    const char* e = getenv("FHE_B200_FUSED_TILE");
    if (!e || atoi(e) == 0) return false;
    return fused_tile_supported(c->plan, c->dnum);
}
// forward transform, pointwise tensor product, inverse transform of  ext [4][B][A][N] (planes 0..3, or 0..1 when squaring)  ->  d [3][B][A][N]
// planes [p0, p0 + np) of ext still need their forward transform when np > 0 (the split path transforms them on other streams)
static int tensor_transform(fhe_b200_bfv* c, uint64_t* ext, uint64_t* d, uint32_t B, bool square, cudaStream_t st) {
    const uint32_t A = c->L + c->R;
    const size_t an = (size_t)A * c->n;
    FusedTile t; t.in = ext; t.out = d; t.limb_begin = 0; t.limb_count = A; t.nb = B; t.square = square;
    for (int p = 0; p < 4; p++) t.in_plane[p] = (size_t)p * B * an;
    for (int q = 0; q < 3; q++) t.out_plane[q] = (size_t)q * B * an;
    t.in_poly = an; t.out_poly = an;
    return launch_fused_tile(c->plan, 0, t, st);
}

static inline uint32_t grid_for(const fhe_b200_bfv* c, size_t items) {
    const size_t w = (items + 255) / 256, cap = (size_t)c->plan->sm_count * 16;
    return (uint32_t)(w < cap ? (w ? w : 1) : cap);
}

// host twin of the hamming-weight ternary sampler
static void host_ternary_hw(std::vector<int8_t>& s, uint32_t n, u64 seed, u64 stream, uint32_t hw, const RngKey& rk) {
    std::vector<uint32_t> perm(n);
    for (uint32_t j = 0; j < n; j++) perm[j] = j;
    s.assign(n, 0);
    if (hw > n) hw = n;
    for (uint32_t i = 0; i < hw; i++) {
        const uint32_t j = i + (uint32_t)(rng_word(rk, seed, stream, i) % (n - i));
        const uint32_t tmp = perm[i]; perm[i] = perm[j]; perm[j] = tmp;
        s[perm[i]] = (rng_word(rk, seed, stream + 1, i) & 1) ? -1 : 1;
    }
}

static uint32_t host_gaussian_cdt(double sigma, uint64_t* cdt, uint32_t cap) {
    uint32_t tail = (uint32_t)std::ceil(6.0 * sigma);
    if (tail + 1 > cap) tail = cap - 1;
    double norm = 1.0;
    for (uint32_t x = 1; x <= tail; x++) norm += 2.0 * std::exp(-((double)x * (double)x) / (2.0 * sigma * sigma));
    double cum = 1.0;
    for (uint32_t k = 0; k <= tail; k++) {
        if (k) cum += 2.0 * std::exp(-((double)k * (double)k) / (2.0 * sigma * sigma));
        const double f = cum / norm;
        cdt[k] = f >= 1.0 ? (1ull << 63) : (uint64_t)(f * 9223372036854775808.0);
    }
    cdt[tail] = 1ull << 63;
    return tail + 1;
}

}  // namespace fhe_b200

using namespace fhe_b200;

extern "C" int fhe_b200_gaussian_cdt(double sigma, uint64_t* h_cdt, uint32_t cap) {
    if (!h_cdt || cap < 2 || !(sigma > 0)) { set_error("gaussian_cdt: bad argument"); return FHE_B200_EINVAL; }
    return (int)host_gaussian_cdt(sigma, h_cdt, cap < (uint32_t)kMaxCdt ? cap : (uint32_t)kMaxCdt);
}

extern "C" int fhe_b200_bfv_destroy(fhe_b200_bfv* c) {
    if (!c) return 0;
    DeviceGuard dev_guard(c->device);
    fhe_b200_lincomb_destroy(c->q2r); fhe_b200_lincomb_destroy(c->scale); fhe_b200_lincomb_destroy(c->r2q);
    fhe_b200_lincomb_destroy(c->moddown); fhe_b200_lincomb_destroy(c->dec);
    for (auto* m : c->modup) fhe_b200_lincomb_destroy(m);
    cudaFree(c->d_consts); cudaFree(c->d_idx);
    cudaFree(c->d_ws);
    for (int i = 0; i < 3; i++) {
        cudaFree(c->d_io[i]);
        if (c->io_stream[i]) cudaStreamDestroy(c->io_stream[i]);
        if (c->io_done[i]) cudaEventDestroy(c->io_done[i]);
    }
    for (int i = 0; i < 2; i++) {
        if (c->mul_stream[i]) cudaStreamDestroy(c->mul_stream[i]);
        if (c->mul_join[i]) cudaEventDestroy(c->mul_join[i]);
    }
    if (c->mul_fork) cudaEventDestroy(c->mul_fork);
    for (int i = 0; i < 3; i++) if (c->mul_ev[i]) cudaEventDestroy(c->mul_ev[i]);
    fhe_b200_plan_destroy(c->plan);
    delete c;
    return 0;
}

extern "C" int fhe_b200_bfv_create(uint32_t n, uint32_t L, uint32_t R, uint32_t K, uint32_t dnum, uint64_t t,
                                   const uint64_t* h_moduli, float sigma, uint32_t hamming_weight, int device, fhe_b200_bfv** out) {
    FHE_REQUIRE(out && h_moduli, "bfv_create: null argument");
    *out = nullptr;
    FHE_REQUIRE(L >= 1 && dnum >= 1 && L % dnum == 0, "bfv_create: L=%u must be a multiple of dnum=%u", L, dnum);
    FHE_REQUIRE(K >= 1 && K <= R, "bfv_create: need 1 <= K <= R (K=%u, R=%u)", K, R);
    FHE_REQUIRE(K >= L / dnum, "bfv_create: special modulus needs K >= alpha = L/dnum limbs (K=%u, alpha=%u)", K, L / dnum);
    FHE_REQUIRE(R >= L + 1, "bfv_create: auxiliary basis needs R >= L+1 limbs so that P_R > t*N*Q (R=%u, L=%u)", R, L);
    FHE_REQUIRE(L + R <= 62, "bfv_create: at most 62 primes in total");
    FHE_REQUIRE(t >= 2 && (t >> 32) == 0, "bfv_create: plaintext modulus must be in [2, 2^32)");
    FHE_REQUIRE(sigma > 0 && sigma < 20, "bfv_create: sigma out of range");
    for (uint32_t i = 0; i < L + R; i++)
        for (uint32_t k = i + 1; k < L + R; k++) FHE_REQUIRE(h_moduli[i] != h_moduli[k], "bfv_create: moduli %u and %u are equal", i, k);
    for (uint32_t i = 0; i < L; i++) FHE_REQUIRE(host::invmod(t % h_moduli[i], h_moduli[i]) != 0, "bfv_create: t is not coprime to q_%u", i);

    DeviceGuard dev_guard(device);                 // the allocations below belong to the context's device
    auto* c = new fhe_b200_bfv();
    c->n = n; c->logn = host::ilog2(n); c->L = L; c->R = R; c->K = K; c->dnum = dnum; c->alpha = L / dnum; c->t = t;
    c->device = device; c->hw = hamming_weight;
    c->primes.assign(h_moduli, h_moduli + L + R);
    int rc = fhe_b200_plan_create(n, h_moduli, L + R, device, &c->plan);
    if (rc) { delete c; return rc; }
    const uint64_t* Q = h_moduli; const uint64_t* P = h_moduli + L;
    const uint32_t W = L + K, alpha = c->alpha;
#define BFV_TRY(e) do { rc = (e); if (rc) { fhe_b200_bfv_destroy(c); return rc; } } while (0)
    BFV_TRY(lincomb_create(make_conv_consts(Q, L, P, R), device, &c->q2r));
    BFV_TRY(lincomb_create(make_scale_consts(Q, L, P, R, t, P, R, true), device, &c->scale));
    BFV_TRY(lincomb_create(make_conv_consts(P, R, Q, L), device, &c->r2q));
    BFV_TRY(lincomb_create(make_conv_consts(P, K, Q, L), device, &c->moddown));
    BFV_TRY(lincomb_create(make_scale_consts(Q, L, nullptr, 0, t, &c->t, 1, false), device, &c->dec));
    // index maps: [K] special limbs | per digit: [alpha] sources, [W-alpha] targets
    std::vector<uint32_t> idx;
    for (uint32_t k = 0; k < K; k++) idx.push_back(L + k);
    std::vector<size_t> off_grp(dnum), off_tgt(dnum);
    for (uint32_t d = 0; d < dnum; d++) {
        std::vector<uint64_t> tm;
        off_grp[d] = idx.size();
        for (uint32_t i = 0; i < alpha; i++) idx.push_back(d * alpha + i);
        off_tgt[d] = idx.size();
        for (uint32_t i = 0; i < W; i++) {
            if (i >= d * alpha && i < (d + 1) * alpha) continue;
            idx.push_back(i); tm.push_back(h_moduli[i]);
        }
        fhe_b200_lincomb* mu = nullptr;
        BFV_TRY(lincomb_create(make_conv_consts(Q + d * alpha, alpha, tm.data(), (uint32_t)tm.size()), device, &mu));
        c->modup.push_back(mu);
    }
    cudaError_t e = cudaMalloc(&c->d_idx, idx.size() * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpy(c->d_idx, idx.data(), idx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    // constants
    std::vector<uint64_t> cst(L + W + L + kMaxCdt + L, 0);
    uint64_t q_mod_t = 1 % t;
    for (uint32_t i = 0; i < L; i++) q_mod_t = host::mulmod(q_mod_t, Q[i] % t, t);
    for (uint32_t i = 0; i < L; i++) {
        const uint64_t qi = Q[i];
        const uint64_t neg = (q_mod_t % qi) ? qi - (q_mod_t % qi) : 0;
        cst[i] = host::mulmod(neg, host::invmod(t % qi, qi), qi);                  // floor(Q/t) mod q_i
        uint64_t pm = 1;
        for (uint32_t k = 0; k < K; k++) pm = host::mulmod(pm, P[k] % qi, qi);
        cst[L + i] = pm;                                                           // P mod q_i (0 for the special limbs)
        cst[L + W + i] = host::invmod(pm, qi);                                     // P^-1 mod q_i
        cst[L + W + L + kMaxCdt + i] = host::shoup(cst[L + W + i], qi);              // its Shoup companion (ModDown epilogue)
    }
    c->h_cdt.resize(kMaxCdt);
    c->cdt_len = host_gaussian_cdt((double)sigma, c->h_cdt.data(), kMaxCdt);
    for (uint32_t k = 0; k < c->cdt_len; k++) cst[L + W + L + k] = c->h_cdt[k];
    if (e == cudaSuccess) e = cudaMalloc(&c->d_consts, cst.size() * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMemcpy(c->d_consts, cst.data(), cst.size() * sizeof(uint64_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("bfv_create: device allocation failed: %s", cudaGetErrorString(e)); fhe_b200_bfv_destroy(c); return FHE_B200_ECUDA; }
    c->d_delta = c->d_consts; c->d_pmodq = c->d_consts + L; c->d_pinv = c->d_consts + L + W; c->d_cdt = c->d_consts + L + W + L; c->d_pinv_s = c->d_cdt + kMaxCdt;
    c->d_idx_p = c->d_idx;
    for (uint32_t d = 0; d < dnum; d++) { c->d_idx_grp.push_back(c->d_idx + off_grp[d]); c->d_idx_tgt.push_back(c->d_idx + off_tgt[d]); }
#undef BFV_TRY
    *out = c;
    return 0;
}

// 256-bit key for the ChaCha20 generator (NULL: back to the reproducible splitmix generator of the tests).  With a key set, the
// 64-bit seeds of keygen / encrypt / ... only separate calls; the randomness is as good as the key (take it from the OS).
extern "C" int fhe_b200_bfv_set_rng_key(fhe_b200_bfv* c, const uint8_t* h_key32) {
    FHE_REQUIRE(c, "bfv_set_rng_key: null context");
    if (!h_key32) { c->rng.on = 0; for (int i = 0; i < 8; i++) c->rng.k[i] = 0; return 0; }
    for (int i = 0; i < 8; i++)
        c->rng.k[i] = (uint32_t)h_key32[4 * i] | ((uint32_t)h_key32[4 * i + 1] << 8) | ((uint32_t)h_key32[4 * i + 2] << 16) | ((uint32_t)h_key32[4 * i + 3] << 24);
    c->rng.on = 1;
    return 0;
}

extern "C" int fhe_b200_bfv_info(const fhe_b200_bfv* c, uint32_t* n, uint32_t* L, uint32_t* R, uint32_t* K, uint32_t* dnum, uint64_t* t) {
    FHE_REQUIRE(c, "bfv_info: null context");
    if (n) *n = c->n; if (L) *L = c->L; if (R) *R = c->R; if (K) *K = c->K; if (dnum) *dnum = c->dnum; if (t) *t = c->t;
    return 0;
}
extern "C" fhe_b200_plan* fhe_b200_bfv_plan(fhe_b200_bfv* c) { return c ? c->plan : nullptr; }

// ---- keys ---------------------------------------------------------------------------------------------------------------
extern "C" int fhe_b200_bfv_keygen(fhe_b200_bfv* c, uint64_t seed_sk, uint64_t seed_pk, uint64_t* d_sk, uint64_t* d_pk, void* stream) {
    FHE_REQUIRE(c && d_sk && d_pk, "bfv_keygen: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard dev_guard(c->device);
    const uint32_t n = c->n, L = c->L, A = c->L + c->R;
    const LimbParams* prm = c->plan->d_params;
    // secret: stream 0 (signs: stream 1 when a hamming weight is set)
    if (c->hw) {
        std::vector<int8_t> s;
        host_ternary_hw(s, n, seed_sk, 0, c->hw, c->rng);
        int8_t* d_s = nullptr;
        FHE_CUDA(cudaMallocAsync(&d_s, n, st));
        FHE_CUDA(cudaMemcpyAsync(d_s, s.data(), n, cudaMemcpyHostToDevice, st));
        FHE_CUDA(cudaStreamSynchronize(st));          // s is a stack-owned host buffer
        expand_small_kernel<<<grid_for(c, n), 256, 0, st>>>(d_sk, d_s, prm, n, 0, A);
        FHE_LAUNCH_CHECK();
        FHE_CUDA(cudaFreeAsync(d_s, st));
    } else {
        sample_small_kernel<0><<<grid_for(c, n), 256, 0, st>>>(d_sk, prm, c->logn, 0, A, 1, seed_sk, -1, 0, c->thr, c->d_cdt, c->cdt_len, c->rng);
        FHE_LAUNCH_CHECK();
    }
    FHE_TRY(launch_ntt(c->plan, d_sk, d_sk, 1, 0, A, false, st));
    // public key: e on stream 2, a on streams 16+i
    uint64_t* pk0 = d_pk; uint64_t* pk1 = d_pk + (size_t)L * n;
    sample_small_kernel<1><<<grid_for(c, n), 256, 0, st>>>(pk0, prm, c->logn, 0, L, 1, seed_pk, -1, 2, 0, c->d_cdt, c->cdt_len, c->rng);
    FHE_LAUNCH_CHECK();
    FHE_TRY(launch_ntt(c->plan, pk0, pk0, 1, 0, L, false, st));
    sample_uniform_kernel<<<grid_for(c, (size_t)L * n), 256, 0, st>>>(pk1, prm, c->logn, 0, L, seed_pk, 16, c->rng);
    FHE_LAUNCH_CHECK();
    pk_finish_kernel<<<grid_for(c, (size_t)L * n), 256, 0, st>>>(pk0, pk1, d_sk, prm, c->logn, (size_t)L * n);
    FHE_LAUNCH_CHECK();
    return 0;
}

extern "C" int fhe_b200_bfv_relinkeygen(fhe_b200_bfv* c, uint64_t seed, const uint64_t* d_sk, uint64_t* d_rlk, void* stream) {
    FHE_REQUIRE(c && d_sk && d_rlk, "bfv_relinkeygen: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard dev_guard(c->device);
    const uint32_t n = c->n, W = c->L + c->K;
    const LimbParams* prm = c->plan->d_params;
    for (uint32_t d = 0; d < c->dnum; d++) {
        uint64_t* b = d_rlk + (size_t)(2 * d) * W * n; uint64_t* a = b + (size_t)W * n;
        const uint64_t base = 1024ull * (d + 1);
        sample_small_kernel<1><<<grid_for(c, n), 256, 0, st>>>(b, prm, c->logn, 0, W, 1, seed, -1, base + 512, 0, c->d_cdt, c->cdt_len, c->rng);
        FHE_LAUNCH_CHECK();
        FHE_TRY(launch_ntt(c->plan, b, b, 1, 0, W, false, st));
        sample_uniform_kernel<<<grid_for(c, (size_t)W * n), 256, 0, st>>>(a, prm, c->logn, 0, W, seed, base, c->rng);
        FHE_LAUNCH_CHECK();
        rlk_finish_kernel<<<grid_for(c, (size_t)W * n), 256, 0, st>>>(b, a, d_sk, nullptr, prm, c->d_pmodq, c->logn, d * c->alpha, (d + 1) * c->alpha, (size_t)W * n);
        FHE_LAUNCH_CHECK();
    }
    return 0;
}

// ---- encrypt / decrypt / add -----------------------------------------------------------------------------------------------
// ---- batch split over two internal streams ------------------------------------------------------------------------------------
// f(first, count, stream) is queued for each half of the batch on its own internal stream, forked from and joined to the caller's
// stream.  Kernels of different character then overlap: the base conversions (tensor pipe) or HBM-bound element-wise passes of one
// half with the IMAD-bound transforms of the other.  FHE_B200_HMULT_STREAMS=1 keeps everything on the caller's stream.
template <class F>
static int fork_join_halves(fhe_b200_bfv* c, uint32_t batch, cudaStream_t st, F&& f) {
    const int env_streams = getenv("FHE_B200_HMULT_STREAMS") ? atoi(getenv("FHE_B200_HMULT_STREAMS")) : 2;      // read at every call
    if (batch < 2 || env_streams < 2) return f(0u, batch, st);
    for (int i = 0; i < 2; i++) {
        if (!c->mul_stream[i]) FHE_CUDA(cudaStreamCreateWithFlags(&c->mul_stream[i], cudaStreamNonBlocking));
        if (!c->mul_join[i]) FHE_CUDA(cudaEventCreateWithFlags(&c->mul_join[i], cudaEventDisableTiming));
    }
    if (!c->mul_fork) FHE_CUDA(cudaEventCreateWithFlags(&c->mul_fork, cudaEventDisableTiming));
    const uint32_t b0 = (batch + 1) / 2, cnt[2] = {b0, batch - b0}, first[2] = {0, b0};
    FHE_CUDA(cudaEventRecord(c->mul_fork, st));
    // from here to the joins nothing returns early: whatever was queued on an internal stream is joined back into the caller's
    int rc = 0;
    bool forked[2] = {false, false};
    for (int i = 0; i < 2 && !rc; i++) {
        if (cudaStreamWaitEvent(c->mul_stream[i], c->mul_fork, 0) != cudaSuccess) { set_error("fork: cudaStreamWaitEvent failed"); rc = FHE_B200_ECUDA; break; }
        forked[i] = true;
        rc = f(first[i], cnt[i], c->mul_stream[i]);
    }
    for (int i = 0; i < 2; i++) {
        if (!forked[i]) continue;
        if (cudaEventRecord(c->mul_join[i], c->mul_stream[i]) != cudaSuccess || cudaStreamWaitEvent(st, c->mul_join[i], 0) != cudaSuccess) {
            if (!rc) { set_error("join: recording or waiting for the internal stream failed"); rc = FHE_B200_ECUDA; }
        }
    }
    return rc;
}

extern "C" int fhe_b200_bfv_encrypt(fhe_b200_bfv* c, uint64_t seed, const uint64_t* d_pt, const uint64_t* d_pk, uint64_t* d_ct,
                                    uint32_t batch, void* stream) {
    FHE_REQUIRE(c && d_pt && d_pk && d_ct, "bfv_encrypt: null argument");
    if (!batch) return 0;
    DeviceGuard dev_guard(c->device);
    const uint32_t n = c->n, L = c->L;
    const LimbParams* prm = c->plan->d_params;
    const size_t ln = (size_t)L * n;
    FHE_TRY(ensure_words(&c->d_ws, &c->ws_words, batch * ln));
    // ciphertext b is batch item b of this call: rng_item_seed(seed, b); a half that starts at polynomial `first` passes first on
    return fork_join_halves(c, batch, (cudaStream_t)stream, [&](uint32_t first, uint32_t cnt, cudaStream_t st) -> int {
        uint64_t* u = c->d_ws + (size_t)first * ln;
        uint64_t* ct = d_ct + (size_t)first * 2 * ln;
        const uint64_t* pt = d_pt + (size_t)first * n;
        sample_small_kernel<0><<<grid_for(c, (size_t)cnt * n), 256, 0, st>>>(u, prm, c->logn, 0, L, cnt, seed, (long long)first, 0, c->thr, c->d_cdt, c->cdt_len, c->rng);
        FHE_LAUNCH_CHECK();
        FHE_TRY(launch_ntt(c->plan, u, u, cnt, 0, L, false, st));
        enc_mul_kernel<<<grid_for(c, cnt * ln), 256, 0, st>>>(ct, u, d_pk, prm, c->logn, L, cnt * ln);
        count_launch();
        FHE_TRY(launch_ntt(c->plan, ct, ct, 2 * cnt, 0, L, true, st));
        enc_finish_kernel<<<grid_for(c, (size_t)cnt * n), 256, 0, st>>>(ct, pt, prm, c->d_delta, c->d_cdt, c->cdt_len, c->logn, L, cnt, seed, (uint64_t)first, c->rng);
        FHE_LAUNCH_CHECK();
        return 0;
    });
}

extern "C" int fhe_b200_bfv_decrypt(fhe_b200_bfv* c, const uint64_t* d_ct, const uint64_t* d_sk, uint64_t* d_pt, uint32_t batch, void* stream) {
    FHE_REQUIRE(c && d_ct && d_sk && d_pt, "bfv_decrypt: null argument");
    if (!batch) return 0;
    DeviceGuard dev_guard(c->device);
    const uint32_t n = c->n, L = c->L;
    const LimbParams* prm = c->plan->d_params;
    const size_t ln = (size_t)L * n;
    FHE_TRY(ensure_words(&c->d_ws, &c->ws_words, batch * ln));
    return fork_join_halves(c, batch, (cudaStream_t)stream, [&](uint32_t first, uint32_t cnt, cudaStream_t st) -> int {
        uint64_t* x = c->d_ws + (size_t)first * ln;
        const uint64_t* ct = d_ct + (size_t)first * 2 * ln;
        // x = NTT(c1) read straight out of the interleaved ciphertexts (no gather copy where the two-pass transform applies),
        // x = INTT(x * s); c0 is added inside the decoding conversion's prologue
        if (c->plan->bal) FHE_TRY(launch_ntt_strided_in(c->plan, x, ct + ln, 2 * ln, cnt, 0, L, false, st));
        else {
            cudaError_t e = cudaMemcpy2DAsync(x, ln * 8, ct + ln, 2 * ln * 8, ln * 8, cnt, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) { set_error("bfv_decrypt: gather failed: %s", cudaGetErrorString(e)); return FHE_B200_ECUDA; }
            FHE_TRY(launch_ntt(c->plan, x, x, cnt, 0, L, false, st));
        }
        dec_mul_kernel<<<grid_for(c, cnt * ln), 256, 0, st>>>(x, d_sk, prm, c->logn, L, cnt * ln); count_launch();
        FHE_TRY(launch_ntt(c->plan, x, x, cnt, 0, L, true, st));
        LcView v; v.in = x; v.in_add = ct; v.in_add_stride = 2 * ln; v.out = d_pt + (size_t)first * n;
        FHE_TRY(lincomb_launch(c->dec, v, n, cnt, st));
        FHE_CUDA(cudaGetLastError());
        return 0;
    });
}

// ---- invariant noise budget (FHEContext::estimate_noise_budget, /root/reference/include/fhe.cuh:142, declared only) ------------
// budget = log2(Q) - log2( || [t (c0 + c1 s)]_Q ||_inf ) - 1 bits (the definition SEAL uses).  The device computes
// w_i = t (c0 + c1 s) mod q_i per limb with the decryption kernels; the host turns every coefficient into mixed-radix digits
// (Garner: d_0 + d_1 q_0 + d_2 q_0 q_1 + ...), which order like the integers themselves: (Q-1)/2 has the digits (q_i-1)/2 and
// Q-1-w the digits q_i-1-d_i, so centring and the magnitude need no multi-precision arithmetic.  A diagnostic: O(N L^2) host work.
extern "C" int fhe_b200_bfv_noise_budget(fhe_b200_bfv* c, const uint64_t* d_ct, const uint64_t* d_sk, uint32_t batch, double* h_bits,
                                         void* stream) {
    FHE_REQUIRE(c && d_ct && d_sk && h_bits, "bfv_noise_budget: null argument");
    if (!batch) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard dev_guard(c->device);
    const uint32_t n = c->n, L = c->L;
    const LimbParams* prm = c->plan->d_params;
    const size_t ln = (size_t)L * n;
    FHE_TRY(ensure_words(&c->d_ws, &c->ws_words, batch * ln));
    uint64_t* x = c->d_ws;
    cudaError_t e = cudaMemcpy2DAsync(x, ln * 8, d_ct + ln, 2 * ln * 8, ln * 8, batch, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { set_error("bfv_noise_budget: gather failed: %s", cudaGetErrorString(e)); return FHE_B200_ECUDA; }
    FHE_TRY(launch_ntt(c->plan, x, x, batch, 0, L, false, st));
    dec_mul_kernel<<<grid_for(c, batch * ln), 256, 0, st>>>(x, d_sk, prm, c->logn, L, batch * ln); count_launch();
    FHE_TRY(launch_ntt(c->plan, x, x, batch, 0, L, true, st));
    dec_add_kernel<<<grid_for(c, batch * ln), 256, 0, st>>>(x, d_ct, prm, c->logn, L, batch * ln); count_launch();
    std::vector<uint64_t> tq(L);
    for (uint32_t i = 0; i < L; i++) tq[i] = c->t % c->primes[i];
    FHE_TRY(fhe_b200_poly_mul_scalar(c->plan, x, x, tq.data(), batch, 0, L, stream));
    std::vector<uint64_t> w(batch * ln);
    FHE_CUDA(cudaMemcpyAsync(w.data(), x, batch * ln * 8, cudaMemcpyDeviceToHost, st));
    FHE_CUDA(cudaStreamSynchronize(st));
    // Garner constants: inv[j][i] = q_j^-1 mod q_i (j < i);  log2 of the partial products
    std::vector<uint64_t> inv((size_t)L * L, 0);
    std::vector<double> lg(L + 1, 0.0);
    for (uint32_t i = 0; i < L; i++) {
        lg[i + 1] = lg[i] + std::log2((double)c->primes[i]);
        for (uint32_t j = 0; j < i; j++) inv[(size_t)j * L + i] = host::invmod(c->primes[j] % c->primes[i], c->primes[i]);
    }
    std::vector<uint64_t> dgt(L);
    for (uint32_t b = 0; b < batch; b++) {
        double worst = -1.0;                                   // log2 of the largest centred magnitude (-1: all zero)
        for (uint32_t k = 0; k < n; k++) {
            for (uint32_t i = 0; i < L; i++) {
                const uint64_t qi = c->primes[i];
                uint64_t u = w[((size_t)b * L + i) * n + k];
                for (uint32_t j = 0; j < i; j++) u = host::mulmod(host::submod(u, dgt[j] % qi, qi), inv[(size_t)j * L + i], qi);
                dgt[i] = u;
            }
            // w > (Q-1)/2 ?  compare digits from the top with (q_i-1)/2
            bool neg = false;
            for (int i = (int)L - 1; i >= 0; i--) {
                const uint64_t h = (c->primes[i] - 1) / 2;
                if (dgt[i] != h) { neg = dgt[i] > h; break; }
            }
            if (neg) {                                         // Q - w = (Q-1-w) + 1
                for (uint32_t i = 0; i < L; i++) dgt[i] = c->primes[i] - 1 - dgt[i];
                for (uint32_t i = 0; i < L; i++) { if (++dgt[i] < c->primes[i]) break; dgt[i] = 0; }
            }
            int top = (int)L - 1;
            while (top >= 0 && dgt[top] == 0) top--;
            if (top < 0) continue;
            const double lead = (double)dgt[top] + (top > 0 ? (double)dgt[top - 1] / (double)c->primes[top - 1] : 0.0);
            const double m = std::log2(lead) + lg[top];
            if (m > worst) worst = m;
        }
        const double bits = worst < 0 ? lg[L] - 1.0 : lg[L] - worst - 1.0;
        h_bits[b] = bits > 0 ? bits : 0.0;
    }
    return 0;
}

extern "C" int fhe_b200_bfv_add(fhe_b200_bfv* c, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out, uint32_t batch, void* stream) {
    FHE_REQUIRE(c && d_a && d_b && d_out, "bfv_add: null argument");
    return launch_elementwise(c->plan, EW_ADD, d_out, d_a, d_b, nullptr, 2 * batch, 0, c->L, (cudaStream_t)stream);
}

extern "C" int fhe_b200_bfv_sub(fhe_b200_bfv* c, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out, uint32_t batch, void* stream) {
    FHE_REQUIRE(c && d_a && d_b && d_out, "bfv_sub: null argument");
    return launch_elementwise(c->plan, EW_SUB, d_out, d_a, d_b, nullptr, 2 * batch, 0, c->L, (cudaStream_t)stream);
}
extern "C" int fhe_b200_bfv_add_plain(fhe_b200_bfv* c, const uint64_t* d_ct, const uint64_t* d_pt, uint64_t* d_out, uint32_t batch,
                                      int subtract, void* stream) {
    FHE_REQUIRE(c && d_ct && d_pt && d_out, "bfv_add_plain: null argument");
    if (!batch) return 0;
    DeviceGuard dev_guard(c->device);
    const size_t total = (size_t)batch * 2 * c->L * c->n;
    add_plain_kernel<<<grid_for(c, total), 256, 0, (cudaStream_t)stream>>>(d_out, d_ct, d_pt, c->plan->d_params, c->d_delta, c->logn, c->L, total, subtract);
    FHE_LAUNCH_CHECK();
    return 0;
}
extern "C" int fhe_b200_bfv_multiply_plain(fhe_b200_bfv* c, const uint64_t* d_ct, const uint64_t* d_pt, uint64_t* d_out, uint32_t batch,
                                           void* stream) {
    FHE_REQUIRE(c && d_ct && d_pt && d_out, "bfv_multiply_plain: null argument");
    if (!batch) return 0;
    DeviceGuard dev_guard(c->device);
    const size_t ln = (size_t)c->L * c->n;
    FHE_TRY(ensure_words(&c->d_ws, &c->ws_words, batch * ln));
    return fork_join_halves(c, batch, (cudaStream_t)stream, [&](uint32_t first, uint32_t cnt, cudaStream_t st) -> int {
        uint64_t* m = c->d_ws + (size_t)first * ln;
        uint64_t* out = d_out + (size_t)first * 2 * ln;
        lift_plain_kernel<<<grid_for(c, cnt * ln), 256, 0, st>>>(m, d_pt + (size_t)first * c->n, c->plan->d_params, c->logn, c->L, cnt * ln);
        FHE_LAUNCH_CHECK();
        FHE_TRY(launch_ntt(c->plan, m, m, cnt, 0, c->L, false, st));
        FHE_TRY(launch_ntt(c->plan, out, d_ct + (size_t)first * 2 * ln, 2 * cnt, 0, c->L, false, st));
        mul_plain_kernel<<<grid_for(c, 2 * cnt * ln), 256, 0, st>>>(out, m, c->plan->d_params, c->logn, c->L, 2 * cnt * ln);
        FHE_LAUNCH_CHECK();
        return launch_ntt(c->plan, out, out, 2 * cnt, 0, c->L, true, st);
    });
}

// ---- hybrid key switching ----------------------------------------------------------------------------------------------------
//   out[b][p] = add_p[b] + ModDown( sum_digits NTT(ModUp(x[b] digit)) * key[digit][p] ),  p = 0, 1   (add_p may be nullptr)
// x: [B] polynomials of L limbs (coefficient form, stride x_stride words), key: [dnum][2][L+K][N] NTT form, out: [B][2][L][N].
// dig [dnum][B][L+K][N] and acc [2][B][L+K][N] are workspace.  Replaces FHEContext::relinearize / key_switch
// (/root/reference/src/fhe.cu:226-235, include/fhe.cuh:134-135; a stub there, intent docs/ARCHITECTURE.md:319-326).
static int key_switch(fhe_b200_bfv* c, const uint64_t* x, size_t x_stride, const uint64_t* d_key, const uint64_t* add0, size_t add0_stride,
                      const uint64_t* add1, size_t add1_stride, uint64_t* d_out, uint32_t B, uint64_t* dig, uint64_t* acc, cudaStream_t st,
                      cudaEvent_t addends_ready = nullptr /* waited for before ModDown reads add0 / add1 */) {
    const uint32_t n = c->n, L = c->L, K = c->K, W = L + K, alpha = c->alpha, dnum = c->dnum;
    const LimbParams* prm = c->plan->d_params;
    const size_t N = n, ln = (size_t)L * N, wn = (size_t)W * N;
    int rc = 0;
    for (uint32_t dg = 0; dg < dnum && !rc; dg++) {
        uint64_t* D = dig + (size_t)dg * B * wn;
        // the digit's own limbs are written through by the conversion kernel (copy_idx = src_idx)
        LcView v; v.in = x; v.in_stride = x_stride; v.src_idx = c->d_idx_grp[dg]; v.out = D; v.out_stride = wn; v.dst_idx = c->d_idx_tgt[dg];
        v.copy_out = D; v.copy_stride = wn;
        rc = lincomb_launch(c->modup[dg], v, n, B, st);
    }
    if (use_fused_tile(c)) {
        // column pass of the digits, then ONE kernel: tile pass of the digits, inner product with the key, inverse tile pass of the sums
        if (!rc) rc = launch_ntt_pass_a(c->plan, dig, dig, dnum * B, 0, W, false, st);
        if (!rc) {
            FusedTile t; t.in = dig; t.out = acc; t.limb_begin = 0; t.limb_count = W; t.nb = B; t.dnum = dnum;
            for (uint32_t d = 0; d < dnum; d++) t.in_plane[d] = (size_t)d * B * wn;
            t.out_plane[0] = 0; t.out_plane[1] = (size_t)B * wn;
            t.in_poly = wn; t.out_poly = wn; t.key = d_key; t.key_poly = wn;
            rc = launch_fused_tile(c->plan, 1, t, st);
        }
        if (!rc) rc = launch_ntt_pass_a(c->plan, acc, acc, 2 * B, 0, W, true, st);
    } else {
        if (!rc) { ScopedLazyForward lazy(true); rc = launch_ntt(c->plan, dig, dig, dnum * B, 0, W, false, st); }    // ks_inner_kernel reduces 128-bit sums
        if (!rc) {
            const uint32_t bpt = ks_inner_bpt(c->plan->sm_count, wn / 2, B);
            const size_t per = (wn / 2) * ((B + bpt - 1) / bpt);
            if (profile_on()) profile_begin(6, B, st);
            ks_inner_kernel<<<grid_for(c, per), 256, 0, st>>>((ulonglong2*)acc, (const ulonglong2*)dig, (const ulonglong2*)d_key, prm, c->logn, 0, W, dnum, B, bpt);
            if (profile_on()) profile_end(st);
            count_launch();
        }
        if (!rc) rc = launch_ntt(c->plan, acc, acc, 2 * B, 0, W, true, st);
    }
    if (addends_ready && cudaStreamWaitEvent(st, addends_ready, 0) != cudaSuccess && !rc) { set_error("key_switch: stream wait failed"); rc = FHE_B200_ECUDA; }
    for (int p = 0; p < 2 && !rc; p++) {
        const uint64_t* s = acc + (size_t)p * B * wn;
        LcView v; v.in = s; v.in_stride = wn; v.src_idx = c->d_idx_p;
        v.sub = s; v.sub_stride = wn; v.epi_scalar = c->d_pinv; v.epi_scalar_shoup = c->d_pinv_s;
        v.add = p ? add1 : add0; v.add_stride = p ? add1_stride : add0_stride;
        v.out = d_out + (size_t)p * ln; v.out_stride = 2 * ln;
        rc = lincomb_launch(c->moddown, v, n, B, st);
    }
    return rc;
}

// ---- multiply, relinearize ---------------------------------------------------------------------------------------------------
// Replaces FHEContext::multiply / relinearize (/root/reference/src/fhe.cu:198-235): tensor over Q u R, exact round(t/Q .), then --
// when d_rlk is given -- hybrid key switching of the third component.  d_scaled (optional) receives the 3-component ciphertext
// [B][3][L][N]; d_out (with d_rlk) the relinearised one [B][2][L][N].  d_a == d_b takes the squaring path: two polynomials are
// extended and transformed instead of four (the tensor kernel reads the same planes for both operands; results are identical).
static size_t multiply_ws_words(const fhe_b200_bfv* c, uint32_t B) {
    const size_t N = c->n, an = (size_t)(c->L + c->R) * N, wn = (size_t)(c->L + c->K) * N, rn = (size_t)c->R * N, ln = (size_t)c->L * N;
    return (size_t)B * (7 * an + 3 * rn + 3 * ln + ((size_t)c->dnum + 2) * wn);
}

static int multiply_half(fhe_b200_bfv* c, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_rlk, uint64_t* d_out,
                         uint64_t* d_scaled, uint32_t batch, uint64_t* ws, cudaStream_t st) {
    const uint32_t n = c->n, L = c->L, R = c->R, K = c->K, A = L + R, W = L + K, dnum = c->dnum, B = batch;
    const LimbParams* prm = c->plan->d_params;
    const size_t N = n, ln = (size_t)L * N, an = (size_t)A * N, wn = (size_t)W * N, rn = (size_t)R * N;
    const bool square = d_a == d_b;
    // workspace: ext [4][B][A][N] | d [3][B][A][N] | sR [3][B][R][N] | sc [3][B][L][N] | dig [dnum][B][W][N] | acc [2][B][W][N]
    const size_t w_ext = 4 * B * an, w_d = 3 * B * an, w_sr = 3 * B * rn, w_sc = 3 * B * ln, w_dig = (size_t)dnum * B * wn, w_acc = 2 * B * wn;
    uint64_t* ext = ws; uint64_t* d = ext + w_ext; uint64_t* sR = d + w_d; uint64_t* sc = sR + w_sr; uint64_t* dig = sc + w_sc; uint64_t* acc = dig + w_dig;
    int rc = 0;
    cudaError_t e = cudaSuccess;
#define STEP(expr) do { if (!rc) rc = (expr); } while (0)
#define COPY2D(dst, dpitch, src, spitch, width, rows) do { if (!rc && e == cudaSuccess) e = cudaMemcpy2DAsync(dst, (dpitch) * 8, src, (spitch) * 8, (width) * 8, rows, cudaMemcpyDeviceToDevice, st); } while (0)
    // 1. the input polynomials into ext: Q limbs copied, R limbs by exact conversion Q -> R
    //    (the conversion kernel also writes the Q limbs through: no separate copy)
    const int planes = square ? 2 : 4;
    for (int p = 0; p < planes; p++) {
        const uint64_t* src = (p < 2 ? d_a : d_b) + (size_t)(p & 1) * ln;           // component p&1 of every ciphertext: stride 2*ln
        uint64_t* dst = ext + (size_t)p * B * an;
        LcView v; v.in = src; v.in_stride = 2 * ln; v.out = dst + ln; v.out_stride = an; v.copy_out = dst; v.copy_stride = an;
        STEP(lincomb_launch(c->q2r, v, n, B, st));
    }
    // 2. NTT over Q u R, 3. tensor, 4. INTT
    if (use_fused_tile(c)) {
        STEP(launch_ntt_pass_a(c->plan, ext, ext, planes * B, 0, A, false, st));
        STEP(tensor_transform(c, ext, d, B, square, st));
        STEP(launch_ntt_pass_a(c->plan, d, d, 3 * B, 0, A, true, st));
    } else {
        { ScopedLazyForward lazy(true); STEP(launch_ntt(c->plan, ext, ext, planes * B, 0, A, false, st)); }               // tensor_kernel reduces 128-bit products
        if (!rc) {
            const size_t per = B * an / 2;
            if (profile_on()) profile_begin(5, B, st);
            tensor_kernel<<<grid_for(c, per), 256, 0, st>>>((ulonglong2*)d, (const ulonglong2*)ext, prm, c->logn, 0, A, per, square ? 0 : 2 * per);
            if (profile_on()) profile_end(st);
            count_launch();
        }
        STEP(launch_ntt(c->plan, d, d, 3 * B, 0, A, true, st));
    }
    // 5. round(t/Q .) in basis R, 6. exact conversion R -> Q
    if (!rc) { LcView v; v.in = d; v.in_stride = an; v.extra = d + ln; v.extra_stride = an; v.out = sR; v.out_stride = rn; rc = lincomb_launch(c->scale, v, n, 3 * B, st); }
    if (!rc) { LcView v; v.in = sR; v.in_stride = rn; v.out = sc; v.out_stride = ln; rc = lincomb_launch(c->r2q, v, n, 3 * B, st); }
    if (d_scaled && !rc) {
        // caller layout [B][3][L][N]; ours is [3][B][L][N]
        for (int p = 0; p < 3; p++) COPY2D(d_scaled + (size_t)p * ln, 3 * ln, sc + (size_t)p * B * ln, ln, ln, B);
    }
    if (e != cudaSuccess) { set_error("bfv_multiply: device copy failed: %s", cudaGetErrorString(e)); return FHE_B200_ECUDA; }
    // 7. relinearise d2 = sc[2]: hybrid key switching, added onto (d0, d1)
    if (d_rlk) STEP(key_switch(c, sc + 2 * B * ln, ln, d_rlk, sc, ln, sc + B * ln, ln, d_out, B, dig, acc, st));
#undef STEP
#undef COPY2D
    if (rc) return rc;
    FHE_CUDA(cudaGetLastError());
    return 0;
}

// One ciphertext pair (batch 1): the multiply itself has two independent branches, run on the caller's stream and on one
// internal stream --
//   inputs : (a0, a1) extended and transformed on the caller's stream, (b0, b1) on the other, joined before the tensor product;
//   outputs: (d0, d1) inverse-transformed, scaled and converted on the caller's stream while d2 goes through the same steps and
//            the key switch up to its ModDown on the other; ModDown adds (d0, d1) and waits for them.
// so that a conversion of one branch overlaps the transforms of the other also when there is no second ciphertext to pair with.
static int multiply_one_split(fhe_b200_bfv* c, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_rlk, uint64_t* d_out,
                              uint64_t* d_scaled, uint64_t* ws, cudaStream_t st) {
    const uint32_t n = c->n, L = c->L, R = c->R, K = c->K, A = L + R, W = L + K, dnum = c->dnum;
    const LimbParams* prm = c->plan->d_params;
    const size_t N = n, ln = (size_t)L * N, an = (size_t)A * N, wn = (size_t)W * N, rn = (size_t)R * N;
    const bool square = d_a == d_b;
    if (!c->mul_stream[1]) FHE_CUDA(cudaStreamCreateWithFlags(&c->mul_stream[1], cudaStreamNonBlocking));
    if (!c->mul_fork) FHE_CUDA(cudaEventCreateWithFlags(&c->mul_fork, cudaEventDisableTiming));
    if (!c->mul_join[1]) FHE_CUDA(cudaEventCreateWithFlags(&c->mul_join[1], cudaEventDisableTiming));
    for (int i = 0; i < 3; i++) if (!c->mul_ev[i]) FHE_CUDA(cudaEventCreateWithFlags(&c->mul_ev[i], cudaEventDisableTiming));
    cudaStream_t sx = c->mul_stream[1];
    uint64_t* ext = ws; uint64_t* d = ext + 4 * an; uint64_t* sR = d + 3 * an; uint64_t* sc = sR + 3 * rn; uint64_t* dig = sc + 3 * ln;
    uint64_t* acc = dig + (size_t)dnum * wn;
    int rc = 0;
    const bool fused = use_fused_tile(c);
    auto extend = [&](int p0, cudaStream_t s) -> int {                 // planes p0, p0 + 1: exact Q -> R (Q limbs written through), forward NTT
        for (int p = p0; p < p0 + 2; p++) {
            const uint64_t* src = (p < 2 ? d_a : d_b) + (size_t)(p & 1) * ln;
            uint64_t* dst = ext + (size_t)p * an;
            LcView v; v.in = src; v.in_stride = 2 * ln; v.out = dst + ln; v.out_stride = an; v.copy_out = dst; v.copy_stride = an;
            FHE_TRY(lincomb_launch(c->q2r, v, n, 1, s));
        }
        if (fused) return launch_ntt_pass_a(c->plan, ext + (size_t)p0 * an, ext + (size_t)p0 * an, 2, 0, A, false, s);
        ScopedLazyForward lazy(true);                                       // tensor_kernel reduces 128-bit products
        return launch_ntt(c->plan, ext + (size_t)p0 * an, ext + (size_t)p0 * an, 2, 0, A, false, s);
    };
    auto descale = [&](uint32_t p0, uint32_t cnt, cudaStream_t s) -> int {      // planes [p0, p0 + cnt) of the tensor: INTT, round(t/Q .), R -> Q
        uint64_t* dp = d + (size_t)p0 * an;
        if (fused) FHE_TRY(launch_ntt_pass_a(c->plan, dp, dp, cnt, 0, A, true, s));
        else FHE_TRY(launch_ntt(c->plan, dp, dp, cnt, 0, A, true, s));
        { LcView v; v.in = dp; v.in_stride = an; v.extra = dp + ln; v.extra_stride = an; v.out = sR + (size_t)p0 * rn; v.out_stride = rn;
          FHE_TRY(lincomb_launch(c->scale, v, n, cnt, s)); }
        { LcView v; v.in = sR + (size_t)p0 * rn; v.in_stride = rn; v.out = sc + (size_t)p0 * ln; v.out_stride = ln;
          FHE_TRY(lincomb_launch(c->r2q, v, n, cnt, s)); }
        if (d_scaled) FHE_CUDA(cudaMemcpyAsync(d_scaled + (size_t)p0 * ln, sc + (size_t)p0 * ln, (size_t)cnt * ln * 8, cudaMemcpyDeviceToDevice, s));
        return 0;
    };
    FHE_CUDA(cudaEventRecord(c->mul_fork, st));
    FHE_CUDA(cudaStreamWaitEvent(sx, c->mul_fork, 0));
    // inputs
    if (!square) { rc = extend(2, sx); cudaEventRecord(c->mul_ev[0], sx); }
    if (!rc) rc = extend(0, st);
    if (!square) cudaStreamWaitEvent(st, c->mul_ev[0], 0);
    if (!rc && fused) rc = tensor_transform(c, ext, d, 1, square, st);
    else if (!rc) {
        const size_t per = an / 2;
        if (profile_on()) profile_begin(5, 1, st);
        tensor_kernel<<<grid_for(c, per), 256, 0, st>>>((ulonglong2*)d, (const ulonglong2*)ext, prm, c->logn, 0, A, per, square ? 0 : 2 * per);
        if (profile_on()) profile_end(st);
        count_launch();
    }
    // outputs
    cudaEventRecord(c->mul_ev[1], st);
    cudaStreamWaitEvent(sx, c->mul_ev[1], 0);
    if (!rc) rc = descale(2, 1, sx);
    if (!rc) rc = descale(0, 2, st);
    cudaEventRecord(c->mul_ev[2], st);                                   // (d0, d1) are in sc
    if (!rc && d_rlk) rc = key_switch(c, sc + 2 * ln, ln, d_rlk, sc, ln, sc + ln, ln, d_out, 1, dig, acc, sx, c->mul_ev[2]);
    cudaEventRecord(c->mul_join[1], sx);
    cudaStreamWaitEvent(st, c->mul_join[1], 0);                          // joined even after an error
    if (rc) return rc;
    FHE_CUDA(cudaGetLastError());
    return 0;
}

// The batch runs as two halves on two internal streams (fork_join_halves): while one half is in a base conversion (tensor pipe,
// 35% of the IMAD pipe) the other is in its transforms (75% of the IMAD pipe), and the hardware interleaves their CTAs -- +4..10% at
// config 4 (tools/gpu_ab_streams.sh).  The halves split the same workspace (its size is linear in the batch).
static int multiply_core(fhe_b200_bfv* c, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_rlk, uint64_t* d_out,
                         uint64_t* d_scaled, uint32_t batch, cudaStream_t st) {
    DeviceGuard dev_guard(c->device);
    FHE_TRY(ensure_words(&c->d_ws, &c->ws_words, multiply_ws_words(c, batch)));
    const size_t ct = 2 * (size_t)c->L * c->n, ct3 = 3 * (size_t)c->L * c->n;
    const int env_streams = getenv("FHE_B200_HMULT_STREAMS") ? atoi(getenv("FHE_B200_HMULT_STREAMS")) : 2;
    if (batch == 1 && env_streams >= 2) return multiply_one_split(c, d_a, d_b, d_rlk, d_out, d_scaled, c->d_ws, st);
    return fork_join_halves(c, batch, st, [&](uint32_t first, uint32_t cnt, cudaStream_t s) -> int {
        const size_t o = first;
        return multiply_half(c, d_a + o * ct, d_b + o * ct, d_rlk, d_out ? d_out + o * ct : nullptr, d_scaled ? d_scaled + o * ct3 : nullptr,
                             cnt, c->d_ws + multiply_ws_words(c, first), s);
    });
}

extern "C" int fhe_b200_bfv_multiply_relin(fhe_b200_bfv* c, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_rlk,
                                           uint64_t* d_out, uint64_t* d_scaled, uint32_t batch, void* stream) {
    FHE_REQUIRE(c && d_a && d_b && d_rlk && d_out, "bfv_multiply_relin: null argument");
    if (!batch) return 0;
    return multiply_core(c, d_a, d_b, d_rlk, d_out, d_scaled, batch, (cudaStream_t)stream);
}

extern "C" int fhe_b200_bfv_multiply(fhe_b200_bfv* c, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out3, uint32_t batch, void* stream) {
    FHE_REQUIRE(c && d_a && d_b && d_out3, "bfv_multiply: null argument");
    if (!batch) return 0;
    return multiply_core(c, d_a, d_b, nullptr, nullptr, d_out3, batch, (cudaStream_t)stream);
}

extern "C" int fhe_b200_bfv_relinearize(fhe_b200_bfv* c, const uint64_t* d_ct3, const uint64_t* d_rlk, uint64_t* d_out, uint32_t batch, void* stream) {
    FHE_REQUIRE(c && d_ct3 && d_rlk && d_out, "bfv_relinearize: null argument");
    FHE_REQUIRE(d_ct3 != d_out, "bfv_relinearize: the output must not alias the input");
    if (!batch) return 0;
    DeviceGuard dev_guard(c->device);
    const size_t ln = (size_t)c->L * c->n, wn = (size_t)(c->L + c->K) * c->n, per = ((size_t)c->dnum + 2) * wn;      // workspace words per ciphertext
    FHE_TRY(ensure_words(&c->d_ws, &c->ws_words, batch * per));
    return fork_join_halves(c, batch, (cudaStream_t)stream, [&](uint32_t first, uint32_t cnt, cudaStream_t st) -> int {
        uint64_t* dig = c->d_ws + (size_t)first * per; uint64_t* acc = dig + (size_t)c->dnum * cnt * wn;
        const uint64_t* in = d_ct3 + (size_t)first * 3 * ln;
        FHE_TRY(key_switch(c, in + 2 * ln, 3 * ln, d_rlk, in, 3 * ln, in + ln, 3 * ln, d_out + (size_t)first * 2 * ln, cnt, dig, acc, st));
        FHE_CUDA(cudaGetLastError());
        return 0;
    });
}

// ---- Galois keys, automorphisms / rotations, modulus chain (declared only in the reference: include/fhe.cuh:59-61,86,109-116) ---
static uint32_t inv_mod_2n(uint32_t g, uint32_t two_n) {          // g odd, two_n a power of two: Newton iteration
    uint32_t x = g;                                                // correct to 3 bits
    for (int i = 0; i < 5; i++) x *= 2u - g * x;
    return x & (two_n - 1);
}

extern "C" int fhe_b200_bfv_galoiskeygen(fhe_b200_bfv* c, uint64_t seed, uint32_t galois_elt, const uint64_t* d_sk, uint64_t* d_gk, void* stream) {
    FHE_REQUIRE(c && d_sk && d_gk, "bfv_galoiskeygen: null argument");
    FHE_REQUIRE((galois_elt & 1u) && galois_elt < 2 * c->n, "bfv_galoiskeygen: the Galois element must be odd and below 2N");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard dev_guard(c->device);
    const uint32_t n = c->n, W = c->L + c->K;
    const LimbParams* prm = c->plan->d_params;
    const size_t wn = (size_t)W * n;
    // s(x^g) in NTT form over the L+K key limbs: INTT(s) -> automorphism -> NTT
    FHE_TRY(ensure_words(&c->d_ws, &c->ws_words, 2 * wn));
    uint64_t* tmp = c->d_ws; uint64_t* sg = tmp + wn;
    FHE_TRY(launch_ntt(c->plan, tmp, d_sk, 1, 0, W, true, st));
    galois_kernel<<<grid_for(c, wn), 256, 0, st>>>(sg, tmp, prm, c->logn, 0, W, inv_mod_2n(galois_elt, 2 * n), wn);
    FHE_LAUNCH_CHECK();
    FHE_TRY(launch_ntt(c->plan, sg, sg, 1, 0, W, false, st));
    for (uint32_t d = 0; d < c->dnum; d++) {
        uint64_t* b = d_gk + (size_t)(2 * d) * wn; uint64_t* a = b + wn;
        const uint64_t base = 1024ull * (d + 1);
        sample_small_kernel<1><<<grid_for(c, n), 256, 0, st>>>(b, prm, c->logn, 0, W, 1, seed, -1, base + 512, 0, c->d_cdt, c->cdt_len, c->rng);
        FHE_LAUNCH_CHECK();
        FHE_TRY(launch_ntt(c->plan, b, b, 1, 0, W, false, st));
        sample_uniform_kernel<<<grid_for(c, wn), 256, 0, st>>>(a, prm, c->logn, 0, W, seed, base, c->rng);
        FHE_LAUNCH_CHECK();
        rlk_finish_kernel<<<grid_for(c, wn), 256, 0, st>>>(b, a, d_sk, sg, prm, c->d_pmodq, c->logn, d * c->alpha, (d + 1) * c->alpha, wn);
        FHE_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int fhe_b200_bfv_apply_galois(fhe_b200_bfv* c, const uint64_t* d_ct, uint32_t galois_elt, const uint64_t* d_gk, uint64_t* d_out,
                                         uint32_t batch, void* stream) {
    FHE_REQUIRE(c && d_ct && d_gk && d_out, "bfv_apply_galois: null argument");
    FHE_REQUIRE((galois_elt & 1u) && galois_elt < 2 * c->n, "bfv_apply_galois: the Galois element must be odd and below 2N");
    if (!batch) return 0;
    DeviceGuard dev_guard(c->device);
    const uint32_t n = c->n, L = c->L, W = L + c->K;
    const size_t ln = (size_t)L * n, wn = (size_t)W * n;
    // workspace per ciphertext: rot [2][L][N] | dig [dnum][W][N] | acc [2][W][N]
    const size_t per = 2 * ln + ((size_t)c->dnum + 2) * wn;
    FHE_TRY(ensure_words(&c->d_ws, &c->ws_words, batch * per));
    const uint32_t ginv = inv_mod_2n(galois_elt, 2 * n);
    return fork_join_halves(c, batch, (cudaStream_t)stream, [&](uint32_t first, uint32_t cnt, cudaStream_t st) -> int {
        const size_t w_rot = 2 * (size_t)cnt * ln, w_dig = (size_t)c->dnum * cnt * wn;
        uint64_t* rot = c->d_ws + (size_t)first * per; uint64_t* dig = rot + w_rot; uint64_t* acc = dig + w_dig;
        galois_kernel<<<grid_for(c, w_rot), 256, 0, st>>>(rot, d_ct + (size_t)first * 2 * ln, c->plan->d_params, c->logn, 0, L, ginv, w_rot);
        FHE_LAUNCH_CHECK();
        // (c0(x^g), 0) + KeySwitch(c1(x^g))
        FHE_TRY(key_switch(c, rot + ln, 2 * ln, d_gk, rot, 2 * ln, nullptr, 0, d_out + (size_t)first * 2 * ln, cnt, dig, acc, st));
        FHE_CUDA(cudaGetLastError());
        return 0;
    });
}

extern "C" int fhe_b200_bfv_mod_switch_to_next(fhe_b200_bfv* c, const uint64_t* d_ct, uint64_t* d_out, uint32_t batch, void* stream) {
    FHE_REQUIRE(c && d_ct && d_out, "bfv_mod_switch_to_next: null argument");
    FHE_REQUIRE(c->L >= 2, "bfv_mod_switch_to_next: the ciphertext modulus has a single limb");
    FHE_REQUIRE(d_out != d_ct, "bfv_mod_switch_to_next: the output is compacted ([..][L-1][N]) and must not alias the input");
    // both components of every ciphertext: [2 batch][L][N] -> [2 batch][L-1][N]
    return fhe_b200_modswitch_drop_last(c->plan, d_out, d_ct, 2 * batch, 0, c->L, stream);
}

// mod_switch_to_level (/root/reference/include/fhe.cuh:110, declared only): the last `drop` limbs of Q go one at a time, each step
// rounding (x - [x]_{q_last}) / q_last exactly like mod_switch_to_next; the result [batch][2][L-drop][N] belongs to the context built
// on the first L - drop limbs.  Intermediate levels ping-pong through the context workspace.
extern "C" int fhe_b200_bfv_mod_switch_to_level(fhe_b200_bfv* c, const uint64_t* d_ct, uint64_t* d_out, uint32_t batch, uint32_t drop,
                                                void* stream) {
    FHE_REQUIRE(c && d_ct && d_out, "bfv_mod_switch_to_level: null argument");
    FHE_REQUIRE(drop < c->L, "bfv_mod_switch_to_level: cannot drop %u of %u limbs", drop, c->L);
    FHE_REQUIRE(d_out != d_ct || drop == 0, "bfv_mod_switch_to_level: the output is compacted and must not alias the input");
    if (!batch) return 0;
    DeviceGuard dev_guard(c->device);
    const size_t pn = (size_t)2 * batch * c->n;                       // words per limb over all components
    if (drop == 0) { if (d_out != d_ct) FHE_CUDA(cudaMemcpyAsync(d_out, d_ct, pn * c->L * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream)); return 0; }
    if (drop > 1) FHE_TRY(ensure_words(&c->d_ws, &c->ws_words, 2 * pn * (c->L - 1)));
    const uint64_t* cur = d_ct;
    for (uint32_t k = 0; k < drop; k++) {
        uint64_t* dst = (k + 1 == drop) ? d_out : c->d_ws + (size_t)(k & 1) * pn * (c->L - 1);
        FHE_TRY(fhe_b200_modswitch_drop_last(c->plan, dst, cur, 2 * batch, 0, c->L - k, stream));
        cur = dst;
    }
    return 0;
}

// Host-buffer entry point.  Chunks of ciphertext pairs rotate over three streams, each with its own device staging buffers:
// the upload of chunk k+1 and the download of chunk k-1 overlap the multiply of chunk k (PCIe is full duplex; with two
// streams the upload of chunk k+2 queued behind the download of chunk k).  The multiplies themselves are serialised through an event because they share the context workspace.
extern "C" int fhe_b200_bfv_multiply_relin_host(fhe_b200_bfv* c, const uint64_t* h_a, const uint64_t* h_b, const uint64_t* d_rlk,
                                                uint64_t* h_out, uint32_t batch) {
    FHE_REQUIRE(c && h_a && h_b && d_rlk && h_out, "bfv_multiply_relin_host: null argument");
    if (!batch) return 0;
    DeviceGuard dev_guard(c->device);
    const size_t ct = 2 * (size_t)c->L * c->n;            // words per ciphertext
    // chunk = G ciphertext pairs (about 64 MiB of input, at least one pair); three staging sets on three streams
    uint32_t G = (uint32_t)(((size_t)64 << 20) / (2 * ct * 8));
    G = G < 1 ? 1 : (G > batch ? batch : G);
    for (int i = 0; i < 3; i++) {
        if (!c->io_stream[i]) FHE_CUDA(cudaStreamCreateWithFlags(&c->io_stream[i], cudaStreamNonBlocking));
        if (!c->io_done[i]) FHE_CUDA(cudaEventCreateWithFlags(&c->io_done[i], cudaEventDisableTiming));
        FHE_TRY(ensure_words(&c->d_io[i], &c->io_words[i], 3 * ct * G));
    }
    // size the shared workspace up front so that no (synchronising) reallocation happens inside the pipeline
    {
        const size_t an = (size_t)(c->L + c->R) * c->n, wn = (size_t)(c->L + c->K) * c->n, rn = (size_t)c->R * c->n, ln = (size_t)c->L * c->n;
        FHE_TRY(ensure_words(&c->d_ws, &c->ws_words, (size_t)G * (7 * an + 3 * rn + 3 * ln + ((size_t)c->dnum + 2) * wn)));
    }
    uint32_t k = 0;
    for (uint32_t i = 0; i < batch; i += G, k++) {
        const uint32_t g = i + G <= batch ? G : batch - i;
        const int s = (int)(k % 3), prev = (int)((k + 2) % 3);
        cudaStream_t st = c->io_stream[s];
        uint64_t* buf = c->d_io[s];
        FHE_CUDA(cudaMemcpyAsync(buf, h_a + i * ct, g * ct * 8, cudaMemcpyHostToDevice, st));
        FHE_CUDA(cudaMemcpyAsync(buf + G * ct, h_b + i * ct, g * ct * 8, cudaMemcpyHostToDevice, st));
        if (k > 0) FHE_CUDA(cudaStreamWaitEvent(st, c->io_done[prev], 0));      // the previous multiply owns the workspace
        FHE_TRY(fhe_b200_bfv_multiply_relin(c, buf, buf + G * ct, d_rlk, buf + 2 * G * ct, nullptr, g, st));
        FHE_CUDA(cudaEventRecord(c->io_done[s], st));
        FHE_CUDA(cudaMemcpyAsync(h_out + i * ct, buf + 2 * G * ct, g * ct * 8, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < 3; i++) FHE_CUDA(cudaStreamSynchronize(c->io_stream[i]));
    return 0;
}

extern "C" int fhe_b200_bfv_tensor(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_ext, uint32_t batch, uint32_t limb_begin,
                                   uint32_t limb_count, void* stream) {
    FHE_REQUIRE(plan && d_out && d_ext, "bfv_tensor: null argument");
    FHE_TRY(check_range(plan, batch, limb_begin, limb_count));
    const size_t per = (size_t)batch * limb_count * plan->n / 2;
    if (!per) return 0;
    DeviceGuard dev_guard(plan->device);
    const size_t w = (per + 255) / 256, cap = (size_t)plan->sm_count * 16;
    tensor_kernel<<<(uint32_t)(w < cap ? w : cap), 256, 0, (cudaStream_t)stream>>>((ulonglong2*)d_out, (const ulonglong2*)d_ext, plan->d_params,
                                                                                 plan->logn, limb_begin, limb_count, per, 2 * per);
    FHE_LAUNCH_CHECK();
    return 0;
}
extern "C" int fhe_b200_bfv_ks_inner(fhe_b200_plan* plan, uint64_t* d_acc, const uint64_t* d_dig, const uint64_t* d_key, uint32_t dnum,
                                     uint32_t batch, uint32_t limb_begin, uint32_t limb_count, void* stream) {
    FHE_REQUIRE(plan && d_acc && d_dig && d_key && dnum >= 1, "bfv_ks_inner: bad argument");
    FHE_TRY(check_range(plan, batch, limb_begin, limb_count));
    if (!((size_t)batch * limb_count)) return 0;
    DeviceGuard dev_guard(plan->device);
    const size_t vpp = (size_t)limb_count * plan->n / 2;
    const uint32_t bpt = ks_inner_bpt(plan->sm_count, vpp, batch);
    const size_t per = vpp * ((batch + bpt - 1) / bpt);
    const size_t w = (per + 255) / 256, cap = (size_t)plan->sm_count * 16;
    ks_inner_kernel<<<(uint32_t)(w < cap ? w : cap), 256, 0, (cudaStream_t)stream>>>((ulonglong2*)d_acc, (const ulonglong2*)d_dig, (const ulonglong2*)d_key,
                                                                                   plan->d_params, plan->logn, limb_begin, limb_count, dnum, batch, bpt);
    FHE_LAUNCH_CHECK();
    return 0;
}
