// RNS linear combination kernel: exact base conversion and t/Q scale-and-round in one primitive.
//
//   z_i   = use_pre ? x_i * pre_i mod s_i : x_i
//   I     = ( sum_i ( z_i * Th_i + hi64(z_i * Tl_i) ) + 2^63 ) >> 64                     (192-bit accumulate)
//   out_k = ( sum_i z_i * M[i][k] + (I mod m_k) * c_k + extra_k * lam_k ) mod m_k      (128-bit lazy accumulate,
//                                                                                        one Barrett reduction)
// Replaces fast_base_conversion_kernel / base_extend (/root/reference/include/rns.cuh:47-49,116-125),
// rns_mod_switch_kernel (:128-136), poly_mod_switch_kernel (include/polynomial.cuh:96-102) and the CRT step of
// decryption (from_rns_crt_kernel, src/rns.cu:117-141).  Integer only: the overflow count / rounding term is a
// 128-bit fixed-point sum, so results are exact and identical on every platform.
//
// One thread per coefficient: the S source residues stay in registers (compile-time SP), the matrix row for the
// current target is read from shared memory as a warp-uniform broadcast.  gridDim.y splits the targets when
// batch*N alone cannot fill the 148 SMs.  IMAD-bound: 4 wide multiplies per (source, target) pair.
#include "lincomb.cuh"
#include "host_math.hpp"
#include <cstring>

namespace fhe_b200 {

struct LcKernelArgs {
    const u64 *src_mod, *pre, *pre_s, *th_hi, *th_lo;
    const u64 *dst_mod, *mu_hi, *mu_lo, *c, *lam, *Mt;
    LcView v;
    uint32_t S, T, logn, k_per_block, use_pre, use_extra, c_is_one;
    const u64* Mt30;     // [T][SP] matrix entries re-split at bit 30: (m & (2^30-1)) | ((m >> 30) << 32)
    size_t total;        // batch * n
};

// 128-bit accumulator += (64-bit v) << sh, sh < 64
__device__ __forceinline__ void add_shifted(u64& hi, u64& lo, u64 v, int sh) {
    add128(hi, lo, sh ? (v >> (64 - sh)) : 0, v << sh);
}
__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }

// One coefficient is handled by TPC adjacent lanes (TPC = 1 or 2): each lane owns SPT of the sources for the prologue and
// the multiply-accumulates, partial sums are exchanged with one shuffle step per target pair, and each lane finishes
// (Barrett, epilogue, store) one target of the pair.  Nothing is computed twice, and at batch 1 the grid has twice the threads.
//
// SPLIT30 (every modulus < 2^60): operands are split at bit 30, so each of the four partial products is below 2^60 and
// sixteen of them fit a 64-bit accumulator.  One (source, target) pair then costs exactly four IMAD.WIDE with the add
// folded into the instruction -- no carry chain, no 64-bit mul.hi emulation (which recomputes the whole product).
template <int SPT, int TPC, bool SPLIT30>
__global__ void __launch_bounds__(256) lincomb_kernel(const LcKernelArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int SP = SPT * TPC;                     // padded source count
    // shared: matrix rows [T][SP], then per-source arrays (5 x SP)
    u64* sMt = sm;
    u64* sSrc = sm + (size_t)a.T * SP;                // src_mod, pre, pre_s, th_hi, th_lo
    const u64* gMt = SPLIT30 ? a.Mt30 : a.Mt;
    for (uint32_t t = threadIdx.x; t < a.T * SP; t += blockDim.x) sMt[t] = gMt[t];
    for (uint32_t t = threadIdx.x; t < (uint32_t)SP; t += blockDim.x) {
        sSrc[t] = a.src_mod[t]; sSrc[SP + t] = a.pre[t]; sSrc[2 * SP + t] = a.pre_s[t];
        sSrc[3 * SP + t] = a.th_hi[t]; sSrc[4 * SP + t] = a.th_lo[t];
    }
    __syncthreads();
    const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t g = gt / TPC;                        // coefficient (over batch * n); a.total is a multiple of 32 / TPC
    const uint32_t sub = (uint32_t)(gt % TPC);
    if (g >= a.total) return;                         // whole warps only (total * TPC is a multiple of 32)
    const uint32_t b = (uint32_t)(g >> a.logn);
    const uint32_t j = (uint32_t)(g & ((1u << a.logn) - 1));
    const size_t nn = (size_t)1 << a.logn;
    const uint32_t s0 = sub * SPT;                    // first source of this lane

    u64 z[SPT];
    u64 f0 = 0, f1 = 0, f2 = 0;                       // 192-bit fixed-point accumulator (this lane's sources)
    const u64* inb = a.v.in + (size_t)b * a.v.in_stride + j;
#pragma unroll
    for (int ii = 0; ii < SPT; ii++) {
        const uint32_t i = s0 + ii;
        u64 x = 0;
        if (i < a.S) {
            x = inb[(size_t)a.v.src_idx[i] * nn];
            if (a.v.in_add) x = add_mod(x, a.v.in_add[(size_t)b * a.v.in_add_stride + j + (size_t)a.v.src_idx[i] * nn], sSrc[i]);
            if (a.v.copy_out) {
                if (a.v.copy_tab) reinterpret_cast<u64*>(a.v.copy_tab[2 * i])[(a.v.copy_poly0 + b) * a.v.copy_tab[2 * i + 1] + j] = x;
                else a.v.copy_out[(size_t)b * a.v.copy_stride + (size_t)a.v.copy_idx[i] * nn + j] = x;
            }
            if (a.use_pre) x = shoup_mul(x, sSrc[SP + i], sSrc[2 * SP + i], sSrc[i]);
        }
        u64 ph, pl;
        mul128(x, sSrc[3 * SP + i], ph, pl);
        const u64 lo = mulhi64(x, sSrc[4 * SP + i]);
        add192(f2, f1, f0, ph, pl);
        add192(f2, f1, f0, 0, lo);
        // SPLIT30: keep the two 30-bit halves side by side in one 64-bit register
        z[ii] = SPLIT30 ? ((x & 0x3fffffffull) | ((x >> 30) << 32)) : x;
    }
    if (TPC == 2) {                                   // combine the two lanes' fixed-point sums
        const u64 o0 = shfl_xor_u64(f0, 1), o1 = shfl_xor_u64(f1, 1), o2 = shfl_xor_u64(f2, 1);
        add192(f2, f1, f0, o1, o0); f2 += o2;
    }
    add192(f2, f1, f0, 0, 1ull << 63);
    const u64 I_hi = f2, I_lo = f1;                   // I = (f2:f1), the rounded integer part

    for (uint32_t kb = 0; kb < a.T; kb += TPC) {
        u64 hi[TPC], lo[TPC];
#pragma unroll
        for (int t = 0; t < TPC; t++) {
            hi[t] = 0; lo[t] = 0;
            const uint32_t k = min(kb + t, a.T - 1);  // a ragged last pair recomputes the last target (never stored twice)
            const u64* row = sMt + (size_t)k * SP + s0;
            if (SPLIT30) {
                u64 u0 = 0, u1 = 0, u2 = 0, u3 = 0;
#pragma unroll
                for (int ii = 0; ii < SPT; ii++) {
                    const u64 mw = row[ii];
                    const u32 z0 = (u32)z[ii], z1 = (u32)(z[ii] >> 32), m0 = (u32)mw, m1 = (u32)(mw >> 32);
                    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(u0) : "r"(z0), "r"(m0));
                    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(u1) : "r"(z0), "r"(m1));
                    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(u2) : "r"(z1), "r"(m0));
                    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(u3) : "r"(z1), "r"(m1));
                    if ((ii & 15) == 15 || ii == SPT - 1) {    // sixteen terms below 2^60 each: flush before overflow
                        add128(hi[t], lo[t], 0, u0); add_shifted(hi[t], lo[t], u1, 30); add_shifted(hi[t], lo[t], u2, 30);
                        add_shifted(hi[t], lo[t], u3, 60);
                        u0 = u1 = u2 = u3 = 0;
                    }
                }
            } else {
#pragma unroll
                for (int ii = 0; ii < SPT; ii++) mac128(hi[t], lo[t], z[ii], row[ii]);
            }
        }
        u64 ah = hi[0], al = lo[0];
        if (TPC == 2) {
            // lane `sub` finishes target kb + sub: it keeps its own partial of that target and receives the partner's
            const u64 sh = sub ? hi[0] : hi[1], sl = sub ? lo[0] : lo[1];
            const u64 rh = shfl_xor_u64(sh, 1), rl = shfl_xor_u64(sl, 1);
            ah = sub ? hi[1] : hi[0]; al = sub ? lo[1] : lo[0];
            add128(ah, al, rh, rl);
        }
        const uint32_t k = kb + sub;
        if (k < a.T) {
            const u64 m = a.dst_mod[k], mh = a.mu_hi[k], ml = a.mu_lo[k];
            // rounded integer term: c_k = 1 for scale-and-round (add I itself), c_k = -Q for conversions (I = overflow count < S)
            if (a.c_is_one) add128(ah, al, I_hi, I_lo); else mac128(ah, al, I_lo, a.c[k]);
            if (a.use_extra) mac128(ah, al, a.v.extra[(size_t)b * a.v.extra_stride + (size_t)a.v.extra_idx[k] * nn + j], a.lam[k]);
            u64 r = barrett128(ah, al, m, mh, ml);
            if (a.v.sub) {
                const size_t eo = (size_t)a.v.epi_idx[k] * nn + j;
                const u64 d = sub_mod(a.v.sub[(size_t)b * a.v.sub_stride + eo], r, m);
                u64 ph, pl;
                mul128(d, a.v.epi_scalar[k], ph, pl);
                r = barrett128(ph, pl, m, mh, ml);
                if (a.v.add) r = add_mod(r, a.v.add[(size_t)b * a.v.add_stride + eo], m);
            }
            if (a.v.out_tab) reinterpret_cast<u64*>(a.v.out_tab[2 * k])[(a.v.out_poly0 + b) * a.v.out_tab[2 * k + 1] + j] = r;
            else a.v.out[(size_t)b * a.v.out_stride + (size_t)a.v.dst_idx[k] * nn + j] = r;
        }
    }
}

// sources per thread the kernel is instantiated for (the source count is padded up to SPT * TPC)
static const int kSupportedSPT[] = {1, 2, 3, 4, 5, 6, 7, 8, 12, 13, 16, 20, 24, 25, 28, 32, 40, 48, 56, 62};

static uint32_t pad_spt(uint32_t need) {
    for (int sp : kSupportedSPT) if ((uint32_t)sp >= need) return (uint32_t)sp;
    return 0;
}

template <int SPT, int TPC>
static int launch_spt(const LcKernelArgs& a, dim3 grid, size_t smem, bool split30, cudaStream_t st) {
    if (split30) {
        if (smem > 48 * 1024) FHE_CUDA(cudaFuncSetAttribute(lincomb_kernel<SPT, TPC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lincomb_kernel<SPT, TPC, true><<<grid, 256, smem, st>>>(a);
    } else {
        if (smem > 48 * 1024) FHE_CUDA(cudaFuncSetAttribute(lincomb_kernel<SPT, TPC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lincomb_kernel<SPT, TPC, false><<<grid, 256, smem, st>>>(a);
    }
    FHE_LAUNCH_CHECK();
    return 0;
}

template <int TPC>
static int launch_tpc(uint32_t spt, const LcKernelArgs& a, dim3 grid, size_t smem, bool split30, cudaStream_t st) {
    switch (spt) {
#define LC_CASE(N_) case N_: return launch_spt<N_, TPC>(a, grid, smem, split30, st);
        LC_CASE(1) LC_CASE(2) LC_CASE(3) LC_CASE(4) LC_CASE(5) LC_CASE(6) LC_CASE(7) LC_CASE(8) LC_CASE(12) LC_CASE(13) LC_CASE(16)
        LC_CASE(20) LC_CASE(24) LC_CASE(25) LC_CASE(28) LC_CASE(32) LC_CASE(40) LC_CASE(48) LC_CASE(56) LC_CASE(62)
#undef LC_CASE
    }
    set_error("lincomb: unsupported sources-per-thread %u", spt);
    return FHE_B200_EINVAL;
}

int lincomb_launch(fhe_b200_lincomb* lc, const LcView& view, uint32_t n, uint32_t batch, cudaStream_t st) {
    FHE_REQUIRE(lc && view.in && (view.out || view.out_tab), "lincomb: null argument");
    FHE_REQUIRE(n >= 32 && (n & (n - 1)) == 0, "lincomb: n_coeffs must be a power of two >= 32");
    FHE_REQUIRE(!lc->use_extra || view.extra, "lincomb: this object needs the extra limbs (scale-and-round)");
    if (!batch) return 0;
    DeviceGuard dev_guard(lc->device);
    LcKernelArgs a;
    a.src_mod = lc->src_mod; a.pre = lc->pre; a.pre_s = lc->pre_s; a.th_hi = lc->th_hi; a.th_lo = lc->th_lo;
    a.dst_mod = lc->dst_mod; a.mu_hi = lc->mu_hi; a.mu_lo = lc->mu_lo; a.c = lc->c; a.lam = lc->lam;
    a.v = view;
    if (!a.v.src_idx) a.v.src_idx = lc->id_src;
    if (!a.v.dst_idx) a.v.dst_idx = lc->id_dst;
    if (!a.v.extra_idx) a.v.extra_idx = lc->id_dst;
    if (!a.v.epi_idx) a.v.epi_idx = lc->id_dst;
    if (!a.v.copy_idx) a.v.copy_idx = a.v.src_idx;
    if (!a.v.copy_stride) a.v.copy_stride = a.v.out_stride ? a.v.out_stride : (size_t)lc->T * n;
    if (!a.v.in_stride) a.v.in_stride = (size_t)lc->S * n;
    if (!a.v.out_stride) a.v.out_stride = (size_t)lc->T * n;
    if (!a.v.extra_stride) a.v.extra_stride = (size_t)lc->T * n;
    if (!a.v.sub_stride) a.v.sub_stride = (size_t)lc->T * n;
    if (!a.v.add_stride) a.v.add_stride = (size_t)lc->T * n;
    a.S = lc->S; a.T = lc->T; a.logn = host::ilog2(n);
    a.use_pre = lc->use_pre; a.use_extra = lc->use_extra; a.c_is_one = lc->use_pre ? 0 : 1;
    a.k_per_block = lc->T;
    a.total = (size_t)batch * n;
    if (!a.v.in_add_stride) a.v.in_add_stride = a.v.in_stride;
    if (lc->use_tc && n % 128 == 0 && !a.v.in_add) {
        const bool prof = profile_on();
        if (prof) profile_begin(4, (uint64_t)batch * lc->S * lc->T, st);
        const int rc = lincomb_tc_launch(lc, a.v, n, batch, st);
        if (prof) profile_end(st);
        return rc;
    }
    if (lc->use_mma && !a.v.in_add) {
        const bool prof = profile_on();
        if (prof) profile_begin(4, (uint64_t)batch * lc->S * lc->T, st);
        const int rc = lincomb_mma_launch(lc, a.v, n, batch, st);
        if (prof) profile_end(st);
        return rc;
    }
    // two lanes per coefficient when the coefficients alone cannot fill the SMs (or when asked to)
    int tpc = (a.total < (size_t)lc->sm_count * 2048 && lc->S >= 2) ? 2 : 1;
    if (lc->force_tpc) tpc = lc->force_tpc;
    if (lc->S < 2) tpc = 1;
    const bool use2 = tpc == 2;
    a.Mt = use2 ? lc->Mt2 : lc->Mt; a.Mt30 = use2 ? lc->Mt30_2 : lc->Mt30;
    // the per-source device arrays are padded to SP1 (one lane) and SP2 = 2 * SPT2 (two lanes): pick the matching set
    if (use2) { a.src_mod = lc->src_mod2; a.pre = lc->pre2; a.pre_s = lc->pre_s2; a.th_hi = lc->th_hi2; a.th_lo = lc->th_lo2; }
    const uint32_t spt = use2 ? lc->SPT2 : lc->SP;
    const uint32_t sp = use2 ? 2 * lc->SPT2 : lc->SP;
    const size_t threads = a.total * (use2 ? 2 : 1);
    const dim3 grid((uint32_t)((threads + 255) / 256), 1);
    const size_t smem = ((size_t)lc->T * sp + 5 * sp) * sizeof(u64);
    FHE_REQUIRE(smem <= 200 * 1024, "lincomb: constant block too large for shared memory");
    const bool prof = profile_on();
    if (prof) profile_begin(4, (uint64_t)batch * lc->S * lc->T, st);          // kind 4: lincomb, units = source-target pairs x batch
    const int rc = use2 ? launch_tpc<2>(spt, a, grid, smem, lc->split30, st) : launch_tpc<1>(spt, a, grid, smem, lc->split30, st);
    if (prof) profile_end(st);
    return rc;
}

int lincomb_create(const LincombConsts& h, int device, fhe_b200_lincomb** out) {
    *out = nullptr;
    FHE_REQUIRE(h.S >= 1 && h.T >= 1, "lincomb: empty basis");
    const uint32_t SP = pad_spt(h.S);
    FHE_REQUIRE(SP != 0, "lincomb: at most 62 source moduli are supported (got %u)", h.S);
    const uint32_t SPT2 = pad_spt((h.S + 1) / 2), SP2 = 2 * SPT2;
    for (uint32_t i = 0; i < h.S; i++) FHE_REQUIRE(h.src_mod[i] > 1 && (h.src_mod[i] >> 61) == 0, "lincomb: source modulus %u out of range", i);
    for (uint32_t k = 0; k < h.T; k++) FHE_REQUIRE(h.dst_mod[k] > 1 && (h.dst_mod[k] >> 61) == 0, "lincomb: target modulus %u out of range", k);
    bool split30 = getenv("FHE_B200_NO_SPLIT30") == nullptr;
    for (uint32_t i = 0; i < h.S; i++) split30 = split30 && (h.src_mod[i] >> 60) == 0;
    for (uint32_t k = 0; k < h.T; k++) split30 = split30 && (h.dst_mod[k] >> 60) == 0;
    DeviceGuard dev_guard(device);
    auto* lc = new fhe_b200_lincomb();
    lc->device = device; lc->S = h.S; lc->T = h.T; lc->SP = SP; lc->SPT2 = SPT2;
    lc->use_pre = h.use_pre; lc->use_extra = h.use_extra; lc->h = h; lc->split30 = split30;
    if (const char* e = getenv("FHE_B200_LINCOMB_TPC")) lc->force_tpc = atoi(e) == 2 ? 2 : (atoi(e) == 1 ? 1 : 0);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) lc->sm_count = prop.multiProcessorCount;
    const uint32_t T = h.T, S = h.S;
    const uint32_t SPmax = SP > SP2 ? SP : SP2;
    // blob (u64 words): for each padding P in {SP, SP2}: 5 x P source arrays, Mt [T][P], Mt30 [T][P];
    //                   5 x T target arrays; identity maps (u32): [SPmax], [T]
    const size_t per_set = [&](uint32_t P) { return 5 * (size_t)P + 2 * (size_t)T * P; }(0);
    (void)per_set;
    const size_t n64 = 5 * (size_t)(SP + SP2) + 2 * (size_t)T * (SP + SP2) + 5 * (size_t)T + (SPmax + 1) / 2 + (T + 1) / 2 + 2;
    std::vector<uint64_t> blob(n64, 0);
    size_t off = 0;
    struct SetOff { size_t src_mod, pre, pre_s, th_hi, th_lo, Mt, Mt30; } so[2];
    const uint32_t pads[2] = {SP, SP2};
    for (int v = 0; v < 2; v++) {
        const uint32_t P = pads[v];
        so[v].src_mod = off; off += P; so[v].pre = off; off += P; so[v].pre_s = off; off += P; so[v].th_hi = off; off += P; so[v].th_lo = off; off += P;
        so[v].Mt = off; off += (size_t)T * P; so[v].Mt30 = off; off += (size_t)T * P;
        for (uint32_t i = 0; i < P; i++) {
            blob[so[v].src_mod + i] = i < S ? h.src_mod[i] : 3;           // padding sources contribute z = 0
            if (i < S) {
                blob[so[v].pre + i] = h.pre[i]; blob[so[v].pre_s + i] = host::shoup(h.pre[i] % h.src_mod[i], h.src_mod[i]);
                blob[so[v].th_hi + i] = h.th_hi[i]; blob[so[v].th_lo + i] = h.th_lo[i];
            }
        }
        for (uint32_t k = 0; k < T; k++)
            for (uint32_t i = 0; i < S; i++) {
                const uint64_t mv = h.M[(size_t)i * T + k];
                blob[so[v].Mt + (size_t)k * P + i] = mv;
                blob[so[v].Mt30 + (size_t)k * P + i] = (mv & 0x3fffffffull) | ((mv >> 30) << 32);
            }
    }
    const size_t o_dst = off; off += T; const size_t o_muh = off; off += T; const size_t o_mul = off; off += T;
    const size_t o_c = off; off += T; const size_t o_lam = off; off += T;
    const size_t o_ids = off; off += (SPmax + 1) / 2; const size_t o_idd = off; off += (T + 1) / 2;
    for (uint32_t k = 0; k < T; k++) {
        blob[o_dst + k] = h.dst_mod[k];
        host::frac128(1, h.dst_mod[k], blob[o_muh + k], blob[o_mul + k]);
        blob[o_c + k] = h.c[k]; blob[o_lam + k] = h.lam[k];
    }
    uint32_t* id_src = reinterpret_cast<uint32_t*>(blob.data() + o_ids);
    uint32_t* id_dst = reinterpret_cast<uint32_t*>(blob.data() + o_idd);
    for (uint32_t i = 0; i < SPmax; i++) id_src[i] = i < S ? i : 0;
    for (uint32_t k = 0; k < T; k++) id_dst[k] = k;
    cudaError_t e = cudaMalloc(&lc->d_blob, n64 * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMemcpy(lc->d_blob, blob.data(), n64 * sizeof(uint64_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("lincomb: device allocation failed: %s", cudaGetErrorString(e)); delete lc; return FHE_B200_ECUDA; }
    const uint64_t* d = lc->d_blob;
    lc->src_mod = d + so[0].src_mod; lc->pre = d + so[0].pre; lc->pre_s = d + so[0].pre_s; lc->th_hi = d + so[0].th_hi; lc->th_lo = d + so[0].th_lo;
    lc->Mt = d + so[0].Mt; lc->Mt30 = d + so[0].Mt30;
    lc->src_mod2 = d + so[1].src_mod; lc->pre2 = d + so[1].pre; lc->pre_s2 = d + so[1].pre_s; lc->th_hi2 = d + so[1].th_hi; lc->th_lo2 = d + so[1].th_lo;
    lc->Mt2 = d + so[1].Mt; lc->Mt30_2 = d + so[1].Mt30;
    lc->dst_mod = d + o_dst; lc->mu_hi = d + o_muh; lc->mu_lo = d + o_mul; lc->c = d + o_c; lc->lam = d + o_lam;
    lc->id_src = reinterpret_cast<const uint32_t*>(d + o_ids);
    lc->id_dst = reinterpret_cast<const uint32_t*>(d + o_idd);
    lc->mma_kt = lincomb_mma_pad_kt(S);
    lc->use_mma = lc->mma_kt != 0 && (size_t)S * T >= 64;
    if (const char* ev = getenv("FHE_B200_LINCOMB_MMA")) lc->use_mma = lc->mma_kt != 0 && atoi(ev) != 0;
    if (lc->use_mma) {
        std::vector<uint2> bf;
        lincomb_mma_build_bfrag(h, lc->mma_kt, bf);
        e = cudaMalloc(&lc->d_bfrag, bf.size() * sizeof(uint2));
        if (e == cudaSuccess) e = cudaMemcpy(lc->d_bfrag, bf.data(), bf.size() * sizeof(uint2), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { set_error("lincomb: device allocation failed: %s", cudaGetErrorString(e)); cudaFree(lc->d_blob); delete lc; return FHE_B200_ECUDA; }
    }
    // tcgen05 path: the byte GEMM with operands in shared memory and accumulators in TMEM (lincomb_tc.cu)
    bool fold = !(getenv("FHE_B200_LINCOMB_TC_FOLD") && atoi(getenv("FHE_B200_LINCOMB_TC_FOLD")) == 0);
    for (uint32_t k = 0; k < T; k++) fold = fold && (h.dst_mod[k] >> 60) == 0 && h.dst_mod[k] > (1ull << 60) - (1ull << 32);
    if (fold && lincomb_tc_smem_bytes(S, T, true) > 220 * 1024) fold = false;
    lc->use_tc = lc->use_mma && lincomb_tc_smem_bytes(S, T, fold) <= 220 * 1024;
    if (const char* ev = getenv("FHE_B200_LINCOMB_TC")) lc->use_tc = atoi(ev) != 0 && lincomb_tc_smem_bytes(S, T, fold) <= 220 * 1024;
    if (lc->use_tc) {
        lc->tc_fold = fold;
        std::vector<uint8_t> bm;
        lincomb_tc_build_b(h, (S + 3) / 4, lc->tc_fold, bm);
        e = cudaMalloc(&lc->d_tc_b, bm.size());
        if (e == cudaSuccess) e = cudaMemcpy(lc->d_tc_b, bm.data(), bm.size(), cudaMemcpyHostToDevice);
        if (lc->tc_fold) {
            std::vector<uint64_t> fc(T);
            for (uint32_t k = 0; k < T; k++) {
                uint64_t hi, lo;
                host::frac128(h.lam[k] % h.dst_mod[k], h.dst_mod[k], hi, lo);      // floor(lam 2^128 / m): the high word is the Shoup companion
                fc[k] = hi;
            }
            if (e == cudaSuccess) e = cudaMalloc(&lc->d_tc_fold, fc.size() * 8);
            if (e == cudaSuccess) e = cudaMemcpy(lc->d_tc_fold, fc.data(), fc.size() * 8, cudaMemcpyHostToDevice);
        }
        if (e != cudaSuccess) { set_error("lincomb: device allocation failed: %s", cudaGetErrorString(e)); cudaFree(lc->d_blob); cudaFree(lc->d_bfrag); delete lc; return FHE_B200_ECUDA; }
    }
    *out = lc;
    return 0;
}

// ---- drop the last limb with rounding ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) modswitch_drop_last_kernel(u64* __restrict__ out, const u64* __restrict__ in,
                                                                  const LimbParams* __restrict__ params, const u64* __restrict__ qlast_inv,
                                                                  uint32_t logn, uint32_t limb_begin, uint32_t limb_count, size_t total) {
    // total = batch * (limb_count-1) * n output elements
    const uint32_t n = 1u << logn;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(g & (n - 1));
        const size_t pl = g >> logn;
        const uint32_t i = (uint32_t)(pl % (limb_count - 1));
        const size_t b = pl / (limb_count - 1);
        const LimbParams P = params[limb_begin + i];
        const u64 ql = params[limb_begin + limb_count - 1].q;
        const u64 r = in[(b * limb_count + (limb_count - 1)) * n + j];
        // centred remainder of the dropped limb, reduced modulo q_i
        u64 rr = barrett128(0, r, P.q, P.mu_hi, P.mu_lo);
        if (r > (ql >> 1)) rr = sub_mod(rr, barrett128(0, ql, P.q, P.mu_hi, P.mu_lo), P.q);
        const u64 d = sub_mod(in[(b * limb_count + i) * n + j], rr, P.q);
        out[g] = mul_mod(d, qlast_inv[i], P);
    }
}

}  // namespace fhe_b200

using namespace fhe_b200;

extern "C" int fhe_b200_lincomb_create_conv(const uint64_t* h_src, uint32_t S, const uint64_t* h_dst, uint32_t T, int device,
                                            fhe_b200_lincomb** out) {
    FHE_REQUIRE(h_src && h_dst && out && S && T, "lincomb_create_conv: null/empty argument");
    return lincomb_create(make_conv_consts(h_src, S, h_dst, T), device, out);
}
extern "C" int fhe_b200_lincomb_create_scale(const uint64_t* h_q, uint32_t L, const uint64_t* h_p, uint32_t R, uint64_t t,
                                             const uint64_t* h_targets, uint32_t T, int with_extra, int device,
                                             fhe_b200_lincomb** out) {
    FHE_REQUIRE(h_q && h_targets && out && L && T, "lincomb_create_scale: null/empty argument");
    FHE_REQUIRE(!with_extra || (h_p && R >= T), "lincomb_create_scale: with_extra needs the targets to be P-basis moduli");
    if (with_extra)
        for (uint32_t k = 0; k < T; k++) {
            bool found = false;
            for (uint32_t r = 0; r < R; r++) found |= (h_p[r] == h_targets[k]);
            FHE_REQUIRE(found, "lincomb_create_scale: target %u is not a modulus of the P basis", k);
        }
    for (uint32_t i = 0; i < L; i++)
        for (uint32_t k = 0; k < T; k++)
            FHE_REQUIRE(h_targets[k] > 1 && host::invmod(h_q[i] % h_targets[k], h_targets[k]) != 0,
                        "lincomb_create_scale: target %u is not coprime to q_%u", k, i);
    return lincomb_create(make_scale_consts(h_q, L, h_p, R, t, h_targets, T, with_extra != 0), device, out);
}
extern "C" int fhe_b200_lincomb_destroy(fhe_b200_lincomb* lc) {
    if (!lc) return 0;
    DeviceGuard dev_guard(lc->device);
    cudaFree(lc->d_blob);
    cudaFree(lc->d_bfrag);
    cudaFree(lc->d_tc_b);
    cudaFree(lc->d_tc_fold);
    delete lc;
    return 0;
}
extern "C" int fhe_b200_lincomb_apply(fhe_b200_lincomb* lc, uint64_t* d_out, const uint64_t* d_in, const uint64_t* d_extra,
                                      uint32_t n_coeffs, uint32_t batch, void* stream) {
    FHE_REQUIRE(lc && d_out && d_in, "lincomb_apply: null argument");
    LcView v; v.in = d_in; v.out = d_out; v.extra = d_extra;
    return lincomb_launch(lc, v, n_coeffs, batch, (cudaStream_t)stream);
}
extern "C" int fhe_b200_lincomb_constants(const fhe_b200_lincomb* lc, uint64_t* pre, uint64_t* th_hi, uint64_t* th_lo,
                                          uint64_t* M, uint64_t* c, uint64_t* lam) {
    FHE_REQUIRE(lc, "lincomb_constants: null handle");
    const auto& h = lc->h;
    if (pre) memcpy(pre, h.pre.data(), h.S * 8);
    if (th_hi) memcpy(th_hi, h.th_hi.data(), h.S * 8);
    if (th_lo) memcpy(th_lo, h.th_lo.data(), h.S * 8);
    if (M) memcpy(M, h.M.data(), (size_t)h.S * h.T * 8);
    if (c) memcpy(c, h.c.data(), h.T * 8);
    if (lam) memcpy(lam, h.lam.data(), h.T * 8);
    return 0;
}
extern "C" int fhe_b200_modswitch_drop_last(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t batch,
                                            uint32_t limb_begin, uint32_t limb_count, void* stream) {
    FHE_REQUIRE(plan && d_out && d_in, "modswitch_drop_last: null argument");
    FHE_TRY(check_range(plan, batch, limb_begin, limb_count));
    FHE_REQUIRE(limb_count >= 2 && limb_count <= 64, "modswitch_drop_last: needs 2..64 limbs");
    FHE_REQUIRE(d_out != d_in, "modswitch_drop_last: the output is compacted ([batch][limbs-1][N]) and must not alias the input");
    const size_t total = (size_t)batch * (limb_count - 1) * plan->n;
    if (!total) return 0;
    DeviceGuard dev_guard(plan->device);
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t h_inv[64];
    const uint64_t ql = plan->moduli[limb_begin + limb_count - 1];
    for (uint32_t i = 0; i + 1 < limb_count; i++) { const uint64_t qi = plan->moduli[limb_begin + i]; h_inv[i] = host::invmod(ql % qi, qi); }
    uint64_t* d_inv = nullptr;
    FHE_CUDA(cudaMallocAsync(&d_inv, 64 * sizeof(uint64_t), st));
    if (cudaMemcpyAsync(d_inv, h_inv, (limb_count - 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st) != cudaSuccess) {
        cudaFreeAsync(d_inv, st);                      // the stream-ordered temporary is released on the error path too
        set_error("modswitch_drop_last: uploading the constants failed");
        return FHE_B200_ECUDA;
    }
    const size_t w = (total + 255) / 256;
    const uint32_t grid = (uint32_t)(w < (size_t)plan->sm_count * 16 ? w : (size_t)plan->sm_count * 16);
    modswitch_drop_last_kernel<<<grid, 256, 0, st>>>(d_out, d_in, plan->d_params, d_inv, plan->logn, limb_begin, limb_count, total);
    FHE_LAUNCH_CHECK();
    FHE_CUDA(cudaFreeAsync(d_inv, st));
    return 0;
}
