// extern "C" entry points: plan lifetime, transforms, element-wise wrappers, host-buffer pipeline.
#include "common.cuh"
#include "tables.hpp"
#include "ntt_bal.cuh"
#include <cstring>
#include <string>
#include <atomic>
#include <mutex>

namespace fhe_b200 {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec { int kind; uint64_t units; cudaEvent_t a, b; };
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
static std::mutex g_prof_mu;
bool profile_on() { return g_prof; }
void profile_begin(int kind, uint64_t units, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r; r.kind = kind; r.units = units;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    g_recs.push_back(r);
}
void profile_end(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_recs.empty()) cudaEventRecord(g_recs.back().b, st);
}

int check_range(const fhe_b200_plan* plan, uint32_t batch, uint32_t limb_begin, uint32_t limb_count) {
    (void)batch;
    FHE_REQUIRE(plan != nullptr, "null plan");
    FHE_REQUIRE((uint64_t)limb_begin + limb_count <= plan->limbs, "limb range [%u, %u) exceeds the plan's %u limbs",
                limb_begin, limb_begin + limb_count, plan->limbs);
    return 0;
}

}  // namespace fhe_b200

using namespace fhe_b200;

extern "C" const char* fhe_b200_last_error(void) { return g_last_error.c_str(); }
extern "C" int fhe_b200_version(void) { return 100; }
extern "C" uint64_t fhe_b200_launch_count(void) { return g_launches.load(); }
extern "C" int fhe_b200_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_recs.clear();
    g_prof = on != 0;
    return 0;
}
extern "C" int fhe_b200_profile_read(int kind, uint64_t* launches, double* total_ms, uint64_t* limb_transforms) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    uint64_t n = 0, u = 0; double ms = 0;
    for (auto& r : g_recs) {
        if (r.kind != kind) continue;
        FHE_CUDA(cudaEventSynchronize(r.b));
        float t = 0; FHE_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
        n++; ms += t; u += r.units;
    }
    if (launches) *launches = n;
    if (total_ms) *total_ms = ms;
    if (limb_transforms) *limb_transforms = u;
    return 0;
}

extern "C" int fhe_b200_plan_create(uint32_t n, const uint64_t* h_moduli, uint32_t n_limbs, int device, fhe_b200_plan** out) {
    FHE_REQUIRE(out != nullptr && h_moduli != nullptr, "plan_create: null argument");
    *out = nullptr;
    FHE_REQUIRE(n >= 512 && n <= 131072 && (n & (n - 1)) == 0, "plan_create: N=%u must be a power of two in [512, 131072]", n);
    FHE_REQUIRE(n_limbs >= 1 && n_limbs <= 4096, "plan_create: bad limb count %u", n_limbs);
    int ndev = 0;
    FHE_CUDA(cudaGetDeviceCount(&ndev));
    FHE_REQUIRE(device >= 0 && device < ndev, "plan_create: device %d not present (%d visible)", device, ndev);
    cudaDeviceProp prop;
    FHE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { set_error("plan_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return FHE_B200_ESTATE; }
    DeviceGuard dev_guard(device);
    {   // keep stream-ordered temporaries (cudaMallocAsync) cached instead of returning them to the OS at every sync
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }

    fhe_b200_plan* p = new (std::nothrow) fhe_b200_plan();
    if (!p) { set_error("out of host memory"); return FHE_B200_ENOMEM; }
    p->n = n; p->logn = host::ilog2(n); p->limbs = n_limbs; p->device = device;
    p->sm_count = prop.multiProcessorCount;
    p->moduli.assign(h_moduli, h_moduli + n_limbs);
    p->hb = lazy_headroom(h_moduli, n_limbs);
    p->near60 = all_near60(h_moduli, n_limbs);
    if (getenv("FHE_B200_NO_NEAR60")) p->near60 = false;
    p->bal = bal_supported((int)p->logn);
    if (const char* e = getenv("FHE_B200_NTT_BAL")) p->bal = p->bal && atoi(e) != 0;
    if (const char* e = getenv("FHE_B200_NTT_CHUNK_MB")) { long mb = atol(e); if (mb > 0) p->chunk_bytes = (size_t)mb << 20; }
    p->h_params.resize(n_limbs);
    std::vector<Twiddle> fwd((size_t)n_limbs * n), inv((size_t)n_limbs * n);
    for (uint32_t l = 0; l < n_limbs; l++) {
        if (build_limb_tables(h_moduli[l], n, fwd.data() + (size_t)l * n, inv.data() + (size_t)l * n, &p->h_params[l])) {
            set_error("plan_create: modulus %llu (limb %u) is not an odd prime < 2^61 with q = 1 mod 2N", (unsigned long long)h_moduli[l], l);
            delete p; return FHE_B200_EINVAL;
        }
    }
    p->tiles = 1u << (p->logn - (uint32_t)tile_lb(p->logn));
    p->p3_entries = tile_p3_entries(p->logn);
    const size_t n12 = (size_t)p->tiles * 256, n3 = (size_t)p->tiles * p->p3_entries;
    std::vector<Twiddle> f12(n_limbs * n12), f3(n_limbs * n3), i12(n_limbs * n12), i3(n_limbs * n3);
    for (uint32_t l = 0; l < n_limbs; l++) {
        build_tile_tables(fwd.data() + (size_t)l * n, p->logn, f12.data() + l * n12, f3.data() + l * n3);
        build_tile_tables(inv.data() + (size_t)l * n, p->logn, i12.data() + l * n12, i3.data() + l * n3);
    }
    const size_t tb = (size_t)n_limbs * n * sizeof(Twiddle);
    cudaError_t e = cudaMalloc(&p->d_fwd, tb);
    auto up = [&](Twiddle** dst, const std::vector<Twiddle>& src) {
        if (e == cudaSuccess) e = cudaMalloc(dst, src.size() * sizeof(Twiddle));
        if (e == cudaSuccess) e = cudaMemcpy(*dst, src.data(), src.size() * sizeof(Twiddle), cudaMemcpyHostToDevice);
    };
    up(&p->d_fwd_p12, f12); up(&p->d_fwd_p3, f3); up(&p->d_inv_p12, i12); up(&p->d_inv_p3, i3);
    if (p->bal) {
        std::vector<Twiddle> fb((size_t)n_limbs * n), ib((size_t)n_limbs * n);
        for (uint32_t l = 0; l < n_limbs; l++) {
            build_bal_tables(fwd.data() + (size_t)l * n, p->logn, fb.data() + (size_t)l * n);
            build_bal_tables(inv.data() + (size_t)l * n, p->logn, ib.data() + (size_t)l * n);
        }
        up(&p->d_fwd_bal, fb); up(&p->d_inv_bal, ib);
    }
    if (e == cudaSuccess) e = cudaMalloc(&p->d_inv, tb);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_params, n_limbs * sizeof(LimbParams));
    if (e == cudaSuccess) e = cudaMemcpy(p->d_fwd, fwd.data(), tb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_inv, inv.data(), tb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_params, p->h_params.data(), n_limbs * sizeof(LimbParams), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("plan_create: device allocation/upload failed: %s", cudaGetErrorString(e));
        fhe_b200_plan_destroy(p);
        return e == cudaErrorMemoryAllocation ? FHE_B200_ENOMEM : FHE_B200_ECUDA;
    }
    *out = p;
    return 0;
}

extern "C" int fhe_b200_plan_destroy(fhe_b200_plan* p) {
    if (!p) return 0;
    DeviceGuard dev_guard(p->device);
    cudaFree(p->d_fwd); cudaFree(p->d_inv); cudaFree(p->d_params);
    cudaFree(p->d_fwd_bal); cudaFree(p->d_inv_bal);
    cudaFree(p->d_fwd_p12); cudaFree(p->d_fwd_p3); cudaFree(p->d_inv_p12); cudaFree(p->d_inv_p3);
    for (int i = 0; i < 3; i++) { if (p->d_stage[i]) cudaFree(p->d_stage[i]); if (p->hs[i]) cudaStreamDestroy(p->hs[i]); }
    delete p;
    return 0;
}
extern "C" uint32_t fhe_b200_plan_n(const fhe_b200_plan* p) { return p ? p->n : 0; }
extern "C" uint32_t fhe_b200_plan_limbs(const fhe_b200_plan* p) { return p ? p->limbs : 0; }
extern "C" int fhe_b200_plan_moduli(const fhe_b200_plan* p, uint64_t* h_out) {
    FHE_REQUIRE(p && h_out, "plan_moduli: null argument");
    memcpy(h_out, p->moduli.data(), p->limbs * sizeof(uint64_t));
    return 0;
}
extern "C" int fhe_b200_plan_tables(const fhe_b200_plan* p, uint32_t limb, uint64_t* h_fwd, uint64_t* h_fwd_shoup,
                                    uint64_t* h_inv, uint64_t* h_inv_shoup) {
    FHE_REQUIRE(p && limb < p->limbs, "plan_tables: bad limb");
    std::vector<Twiddle> t(p->n);
    FHE_CUDA(cudaMemcpy(t.data(), p->d_fwd + (size_t)limb * p->n, p->n * sizeof(Twiddle), cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < p->n; i++) { if (h_fwd) h_fwd[i] = t[i].w; if (h_fwd_shoup) h_fwd_shoup[i] = t[i].ws; }
    FHE_CUDA(cudaMemcpy(t.data(), p->d_inv + (size_t)limb * p->n, p->n * sizeof(Twiddle), cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < p->n; i++) { if (h_inv) h_inv[i] = t[i].w; if (h_inv_shoup) h_inv_shoup[i] = t[i].ws; }
    return 0;
}

extern "C" int fhe_b200_ntt_forward(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t batch,
                                    uint32_t limb_begin, uint32_t limb_count, void* stream) {
    FHE_REQUIRE(plan && d_out && d_in, "ntt_forward: null argument");
    return launch_ntt(plan, d_out, d_in, batch, limb_begin, limb_count, false, (cudaStream_t)stream);
}
extern "C" int fhe_b200_ntt_inverse(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_in, uint32_t batch,
                                    uint32_t limb_begin, uint32_t limb_count, void* stream) {
    FHE_REQUIRE(plan && d_out && d_in, "ntt_inverse: null argument");
    return launch_ntt(plan, d_out, d_in, batch, limb_begin, limb_count, true, (cudaStream_t)stream);
}

#define EW_ENTRY(NAME, OP)                                                                                          \
    extern "C" int NAME(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_a, const uint64_t* d_b, uint32_t batch, \
                        uint32_t limb_begin, uint32_t limb_count, void* stream) {                                  \
        FHE_REQUIRE(plan && d_out && d_a && d_b, #NAME ": null argument");                                          \
        return launch_elementwise(plan, OP, d_out, d_a, d_b, nullptr, batch, limb_begin, limb_count, (cudaStream_t)stream); \
    }
EW_ENTRY(fhe_b200_poly_add, EW_ADD)
EW_ENTRY(fhe_b200_poly_sub, EW_SUB)
EW_ENTRY(fhe_b200_poly_mul, EW_MUL)
EW_ENTRY(fhe_b200_poly_mul_scalar, EW_MUL_SCALAR)
EW_ENTRY(fhe_b200_poly_add_scalar, EW_ADD_SCALAR)
#undef EW_ENTRY

extern "C" int fhe_b200_poly_mac(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_acc, const uint64_t* d_a,
                                 const uint64_t* d_b, uint32_t batch, uint32_t limb_begin, uint32_t limb_count, void* stream) {
    FHE_REQUIRE(plan && d_out && d_acc && d_a && d_b, "poly_mac: null argument");
    return launch_elementwise(plan, EW_MAC, d_out, d_a, d_b, d_acc, batch, limb_begin, limb_count, (cudaStream_t)stream);
}
extern "C" int fhe_b200_poly_negate(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_a, uint32_t batch,
                                    uint32_t limb_begin, uint32_t limb_count, void* stream) {
    FHE_REQUIRE(plan && d_out && d_a, "poly_negate: null argument");
    return launch_elementwise(plan, EW_NEG, d_out, d_a, nullptr, nullptr, batch, limb_begin, limb_count, (cudaStream_t)stream);
}

extern "C" int fhe_b200_negacyclic_mul(fhe_b200_plan* plan, uint64_t* d_out, const uint64_t* d_a, const uint64_t* d_b,
                                       uint32_t batch, uint32_t limb_begin, uint32_t limb_count, void* stream) {
    FHE_REQUIRE(plan && d_out && d_a && d_b, "negacyclic_mul: null argument");
    FHE_TRY(check_range(plan, batch, limb_begin, limb_count));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = (size_t)batch * limb_count * plan->n * sizeof(uint64_t);
    if (!bytes) return 0;
    DeviceGuard dev_guard(plan->device);
    // N <= 4096: one fused kernel (both forward transforms, the pointwise product and the inverse in shared memory) while the
    // grid is small enough that its single CTA per SM does not cost throughput; FHE_B200_MUL_FUSED = 0 | 1 overrides
    if (plan->logn <= 12) {
        const int env = getenv("FHE_B200_MUL_FUSED") ? atoi(getenv("FHE_B200_MUL_FUSED")) : -1;
        const bool fused = env >= 0 ? env != 0 : (size_t)batch * limb_count <= (size_t)4 * plan->sm_count;
        if (fused) return launch_negacyclic_mul_fused(plan, d_out, d_a, d_b, batch, limb_begin, limb_count, st);
    }
    uint64_t* tmp = nullptr;                       // NTT(b); NTT(a) goes straight into d_out
    FHE_CUDA(cudaMallocAsync(&tmp, bytes, st));
    int rc = launch_ntt(plan, tmp, d_b, batch, limb_begin, limb_count, false, st);
    if (!rc) rc = launch_ntt(plan, d_out, d_a, batch, limb_begin, limb_count, false, st);
    if (!rc) rc = launch_elementwise(plan, EW_MUL, d_out, d_out, tmp, nullptr, batch, limb_begin, limb_count, st);
    if (!rc) rc = launch_ntt(plan, d_out, d_out, batch, limb_begin, limb_count, true, st);
    cudaFreeAsync(tmp, st);
    return rc;
}

// Host-buffer path: three staging buffers, three streams; chunk c is uploaded on stream c%3, transformed there and
// downloaded there, so the upload of chunk c+1 and the download of chunk c-1 overlap the transform of chunk c
// (PCIe is full duplex).  Chunks are whole polynomials' limb groups.
extern "C" int fhe_b200_ntt_host(fhe_b200_plan* plan, uint64_t* h_data, uint32_t batch, uint32_t limb_begin,
                                 uint32_t limb_count, int direction) {
    FHE_REQUIRE(plan && h_data, "ntt_host: null argument");
    FHE_REQUIRE(direction >= 0 && direction <= 2, "ntt_host: direction must be 0, 1 or 2");
    FHE_TRY(check_range(plan, batch, limb_begin, limb_count));
    if (!batch || !limb_count) return 0;
    DeviceGuard dev_guard(plan->device);
    const size_t poly_bytes = (size_t)limb_count * plan->n * sizeof(uint64_t);
    // chunk size 64 MiB (measured on the B200 box, config 3 end to end: 64 MiB 172 k, 32 MiB 171 k, 16 MiB 161 k, 8 MiB 161 k
    // limb-transforms/s -- the link sustains ~45 GB/s per direction under duplex load whatever the chunk); FHE_B200_HOST_CHUNK_MB overrides
    size_t chunk_bytes = (size_t)64 << 20;
    if (const char* e = getenv("FHE_B200_HOST_CHUNK_MB")) { const long mb = atol(e); if (mb > 0) chunk_bytes = (size_t)mb << 20; }
    size_t polys_per_chunk = chunk_bytes / poly_bytes; if (!polys_per_chunk) polys_per_chunk = 1;
    const size_t need = polys_per_chunk * poly_bytes;
    if (plan->stage_bytes < need) {
        plan->stage_bytes = 0;                          // nothing usable until all three buffers exist again (an allocation may fail midway)
        for (int i = 0; i < 3; i++) {
            if (plan->d_stage[i]) { cudaFree(plan->d_stage[i]); plan->d_stage[i] = nullptr; }
            if (!plan->hs[i]) FHE_CUDA(cudaStreamCreateWithFlags(&plan->hs[i], cudaStreamNonBlocking));
            FHE_CUDA(cudaMalloc(&plan->d_stage[i], need));
        }
        plan->stage_bytes = need;
    }
    int c = 0;
    for (size_t b0 = 0; b0 < batch; b0 += polys_per_chunk, c++) {
        const uint32_t nb = (uint32_t)((b0 + polys_per_chunk <= batch) ? polys_per_chunk : batch - b0);
        const int s = c % 3;
        uint64_t* h = h_data + b0 * (size_t)limb_count * plan->n;
        FHE_CUDA(cudaMemcpyAsync(plan->d_stage[s], h, nb * poly_bytes, cudaMemcpyHostToDevice, plan->hs[s]));
        if (direction == 0 || direction == 2) FHE_TRY(launch_ntt(plan, plan->d_stage[s], plan->d_stage[s], nb, limb_begin, limb_count, false, plan->hs[s]));
        if (direction == 1 || direction == 2) FHE_TRY(launch_ntt(plan, plan->d_stage[s], plan->d_stage[s], nb, limb_begin, limb_count, true, plan->hs[s]));
        FHE_CUDA(cudaMemcpyAsync(h, plan->d_stage[s], nb * poly_bytes, cudaMemcpyDeviceToHost, plan->hs[s]));
    }
    for (int i = 0; i < 3; i++) FHE_CUDA(cudaStreamSynchronize(plan->hs[i]));
    return 0;
}
