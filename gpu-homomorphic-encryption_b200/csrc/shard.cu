// Limb-sharded BFV multiply + relinearize over G GPUs of one NVLink / NVSwitch domain (BASELINE.json config 4).
//
// The reference only sketches this ("RNS limbs distributed over GPUs, NVLink for the exchange", /root/reference/docs/ARCHITECTURE.md:501-512,
// README.md:320); the operation sharded here is FHEContext::multiply + relinearize (/root/reference/src/fhe.cu:199-235).
//
// A multiply alternates between two kinds of work:
//   * per-LIMB work   -- NTTs, the tensor product, the key-switch inner product: every limb on its own, all N coefficients;
//   * per-COEFFICIENT work -- exact RNS base conversions / scale-and-round: every coefficient on its own, all limbs.
// So the data lives in two layouts and moves between them four times per multiply:
//   limb-shard:        rank g owns a block of limbs of every polynomial, all N coefficients       (NTT domain work)
//   coefficient-shard: rank g owns coefficients [g N/G, (g+1) N/G) of every limb of every polynomial (conversions)
// Each transposition moves 1/G of what an all-gather of the source limbs would, and it is not a separate step:
// the PRODUCING kernel stores straight into the consumer GPU's buffer over NVLink (peer pointers from CUDA IPC) --
//   phase 0  Q -> R extension of the inputs (conversion kernel): target limb k goes to the GPU that owns limb k    (scatter tables)
//   phase 1  last pass of the inverse NTT of the tensor product: coefficient block h goes to GPU h                 (BalScatter)
//   phase 2  ModUp of the key-switch digits (conversion kernel): as phase 0 over the L+K key limbs
//   phase 3  last pass of the inverse NTT of the key-switch accumulators: as phase 1
// so the transfer overlaps the arithmetic of the same kernel warp by warp.  Ordering between GPUs is by epoch flags: after its
// producing kernels a rank stores the epoch of the operation into flag[phase][rank] of every peer (release, system scope); a
// one-warp kernel in the consumer's stream spins on its own flags (acquire) before the consuming kernels start.  There is no
// host synchronisation and no NCCL call on the data path; the host language only carries the 128-byte handles between the
// ranks once (torch.distributed all_gather_object / MPI / a pipe).  A receive buffer of phase p is rewritten by a peer only
// after that peer has passed a wait that this rank satisfies after its last read of the buffer (every phase has its own buffer,
// see DESIGN.md section 6), so operations can be issued back to back without a barrier.
//
// Layouts (uint64).  Nc = N/G, A = L+R, W = L+K, cA/cW = limbs of A / W this rank owns.
//   sharded ciphertext   [B][2][L][Nc]      coefficient form, coefficients [rank Nc, (rank+1) Nc)
//   key slice            [dnum][2][cW][N]   NTT form, limbs w_begin .. w_begin + cW of the relinearisation key
//   EXT  (recv, phase 0) [2 operands][B][2][cA][N]        D2   (recv, phase 1) [3B][A][Nc]
//   DIG  (recv, phase 2) [B][dnum][cW][N]                 ACC2 (recv, phase 3) [2B][W][Nc]
#include "bfv.cuh"
#include "host_math.hpp"
#include <cstring>
#include <unistd.h>

namespace fhe_b200 {

constexpr int kMaxRanks = 16;
constexpr int kPhases = 4;
constexpr uint32_t kHandleMagic = 0x53484232u;       // "SHB2"

struct ShardHandle {                                  // what fhe_b200_shard_handle exports (FHE_B200_SHARD_HANDLE_BYTES = 128)
    uint32_t magic; int32_t pid; int32_t device; uint32_t world;
    uint64_t ptr; uint64_t slab_words;
    cudaIpcMemHandle_t ipc;
    uint8_t pad[128 - 32 - sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(ShardHandle) == 128, "handle size is part of the C ABI");

// control block at the start of every slab
struct ShardCtl {
    unsigned long long flag[kPhases][kMaxRanks];      // flag[p][r] = epoch of the last operation whose phase p rank r has delivered
    unsigned long long epoch;                         // operations started by THIS rank
    unsigned int err;                                 // 0, or 1 + the rank a wait gave up on
    unsigned int pad;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct PeerCtl { ShardCtl* p[kMaxRanks]; };

// runs after the producing kernels of `phase` in the same stream: their stores (local and peer) are complete; publish the epoch
__global__ void shard_signal_kernel(ShardCtl* mine, const PeerCtl peers, int phase, int rank, int world, int bump) {
    __shared__ unsigned long long e;
    if (threadIdx.x == 0) { e = mine->epoch + (bump ? 1ull : 0ull); if (bump) mine->epoch = e; }
    __syncthreads();
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(&peers.p[threadIdx.x]->flag[phase][rank], e);
    }
}
// runs before the consuming kernels: every rank must have delivered `phase` of the current operation
__global__ void shard_wait_kernel(ShardCtl* mine, int phase, int world, unsigned long long timeout_ns) {
    if ((int)threadIdx.x >= world) return;
    const unsigned long long want = mine->epoch;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(&mine->flag[phase][threadIdx.x]) < want) {
        if (global_ns() - t0 > timeout_ns) { atomicCAS(&mine->err, 0u, 1u + threadIdx.x); break; }
        __nanosleep(40);
    }
    __threadfence_system();
}

// ext [2 operands][B][2][cnt][N] (NTT form) -> d [B][3][cnt][N]:  d0 = a0 b0, d1 = a0 b1 + a1 b0, d2 = a1 b1; limb l = plan limb limb_begin + l
__global__ void __launch_bounds__(256) shard_tensor_kernel(ulonglong2* __restrict__ d, const ulonglong2* __restrict__ ext,
                                                           const LimbParams* __restrict__ params, uint32_t logn, uint32_t limb_begin,
                                                           uint32_t cnt, uint32_t B, size_t b_operand /* vectors: B*2*cnt*N/2, 0 when squaring */) {
    const size_t pv = ((size_t)cnt << logn) / 2;                 // vectors per polynomial
    const size_t total = pv * B;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
        const size_t b = v / pv, r = v % pv;
        const LimbParams P = params[limb_begin + (uint32_t)((2 * r) >> logn)];
        const ulonglong2* ea = ext + (2 * b) * pv + r;
        const ulonglong2* eb = ea + b_operand;
        const ulonglong2 a0 = ea[0], a1 = ea[pv], b0 = eb[0], b1 = eb[pv];
        ulonglong2 r0, r1, r2;
        u64 hi, lo;
        r0.x = mul_mod(a0.x, b0.x, P); r0.y = mul_mod(a0.y, b0.y, P);
        r2.x = mul_mod(a1.x, b1.x, P); r2.y = mul_mod(a1.y, b1.y, P);
        hi = 0; lo = 0; mac128(hi, lo, a0.x, b1.x); mac128(hi, lo, a1.x, b0.x); r1.x = barrett128(hi, lo, P.q, P.mu_hi, P.mu_lo);
        hi = 0; lo = 0; mac128(hi, lo, a0.y, b1.y); mac128(hi, lo, a1.y, b0.y); r1.y = barrett128(hi, lo, P.q, P.mu_hi, P.mu_lo);
        ulonglong2* o = d + (3 * b) * pv + r;
        o[0] = r0; o[pv] = r1; o[2 * pv] = r2;
    }
}
// dig [B][dnum][cnt][N], key [dnum][2][cnt][N] (this rank's slice) -> acc [B][2][cnt][N]
__global__ void __launch_bounds__(256) shard_ks_inner_kernel(ulonglong2* __restrict__ acc, const ulonglong2* __restrict__ dig,
                                                             const ulonglong2* __restrict__ key, const LimbParams* __restrict__ params,
                                                             uint32_t logn, uint32_t limb_begin, uint32_t cnt, uint32_t dnum, uint32_t B) {
    const size_t pv = ((size_t)cnt << logn) / 2;
    const size_t total = pv * B;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
        const size_t b = v / pv, r = v % pv;
        const LimbParams P = params[limb_begin + (uint32_t)((2 * r) >> logn)];
        u64 h0x = 0, l0x = 0, h0y = 0, l0y = 0, h1x = 0, l1x = 0, h1y = 0, l1y = 0;
        for (uint32_t dg = 0; dg < dnum; dg++) {
            const ulonglong2 x = dig[(b * dnum + dg) * pv + r];
            const ulonglong2 kb = key[(size_t)(2 * dg) * pv + r], ka = key[(size_t)(2 * dg + 1) * pv + r];
            mac128(h0x, l0x, x.x, kb.x); mac128(h0y, l0y, x.y, kb.y);
            mac128(h1x, l1x, x.x, ka.x); mac128(h1y, l1y, x.y, ka.y);
        }
        ulonglong2 o0, o1;
        o0.x = barrett128(h0x, l0x, P.q, P.mu_hi, P.mu_lo); o0.y = barrett128(h0y, l0y, P.q, P.mu_hi, P.mu_lo);
        o1.x = barrett128(h1x, l1x, P.q, P.mu_hi, P.mu_lo); o1.y = barrett128(h1y, l1y, P.q, P.mu_hi, P.mu_lo);
        acc[(2 * b) * pv + r] = o0; acc[(2 * b + 1) * pv + r] = o1;
    }
}

// contiguous block partition; the first (total % world) ranks get one extra item
static void block_partition(uint32_t total, uint32_t world, uint32_t* begin /* [world+1] */) {
    const uint32_t base = total / world, rem = total % world;
    begin[0] = 0;
    for (uint32_t r = 0; r < world; r++) begin[r + 1] = begin[r] + base + (r < rem ? 1 : 0);
}

}  // namespace fhe_b200

using namespace fhe_b200;

struct fhe_b200_shard {
    fhe_b200_bfv* ctx = nullptr;
    int rank = 0, world = 1, logw = 0;
    uint32_t max_batch = 1, nc = 0;
    uint32_t a_begin[kMaxRanks + 1] = {0}, w_begin[kMaxRanks + 1] = {0};
    uint32_t ca_max = 0, cw_max = 0;
    // the slab: control block + the four receive buffers, identical layout on every rank (sized for the largest limb block)
    uint64_t* slab = nullptr; size_t slab_words = 0;
    size_t off_ext = 0, off_d2 = 0, off_dig = 0, off_acc2 = 0;              // in words from the slab base
    uint64_t* peer[kMaxRanks] = {nullptr};                                    // slab bases (peer[rank] = own slab)
    bool opened[kMaxRanks] = {false};                                         // opened through CUDA IPC (to be closed)
    bool connected = false;
    // local workspace: d [B][3][cA][N] | sR [3B][R][Nc] | sc [3B][L][Nc] | acc [B][2][cW][N]
    uint64_t* ws = nullptr; size_t off_d = 0, off_sr = 0, off_sc = 0, off_acc = 0;
    // scatter tables on the device: q2r out [R] / copy [L];  per digit: out [W-alpha] / copy [alpha]   (pairs {address, words per polynomial})
    uint64_t* d_tabs = nullptr;
    const uint64_t *tab_q2r_out = nullptr, *tab_q2r_copy = nullptr;
    std::vector<const uint64_t*> tab_up_out, tab_up_copy;
    unsigned long long timeout_ns = 10ull * 1000 * 1000 * 1000;
    // CUDA graphs of whole multiplies, keyed by the argument tuple (the flags are device-side epochs, so a graph replays correctly):
    // at one ciphertext pair per multiply a rank's 27 kernels are 5-15 us each and the launch gaps are a fifth of the time
    struct GraphKey { const void *a, *b, *key; void* out; uint32_t batch; int fused; };
    struct GraphEntry { GraphKey k; cudaGraphExec_t exec; uint32_t seen; };
    std::vector<GraphEntry> graphs;
    uint64_t nvlink_words_per_op = 0;                                         // words this rank stores into OTHER ranks per multiply (batch 1)
};

static ShardCtl* ctl_of(uint64_t* slab) { return reinterpret_cast<ShardCtl*>(slab); }
constexpr size_t kCtlWords = (sizeof(ShardCtl) + 255) / 256 * 32;            // control block, padded to 256 bytes

extern "C" int fhe_b200_shard_partition(uint32_t total, uint32_t rank, uint32_t world, uint32_t* begin, uint32_t* count) {
    FHE_REQUIRE(world >= 1 && world <= (uint32_t)kMaxRanks && rank < world, "shard_partition: bad rank %u / world %u", rank, world);
    uint32_t b[kMaxRanks + 1];
    block_partition(total, world, b);
    if (begin) *begin = b[rank];
    if (count) *count = b[rank + 1] - b[rank];
    return 0;
}

extern "C" int fhe_b200_shard_destroy(fhe_b200_shard* s) {
    if (!s) return 0;
    DeviceGuard dev_guard(s->ctx->device);
    cudaDeviceSynchronize();
    for (auto& g : s->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (int r = 0; r < s->world; r++) if (s->opened[r] && s->peer[r]) cudaIpcCloseMemHandle(s->peer[r]);
    cudaFree(s->slab); cudaFree(s->ws); cudaFree(s->d_tabs);
    delete s;
    return 0;
}

extern "C" int fhe_b200_shard_create(fhe_b200_bfv* c, int rank, int world, uint32_t max_batch, fhe_b200_shard** out) {
    FHE_REQUIRE(c && out, "shard_create: null argument");
    *out = nullptr;
    FHE_REQUIRE(world >= 1 && world <= kMaxRanks && (world & (world - 1)) == 0, "shard_create: world size %d must be a power of two <= %d", world, kMaxRanks);
    FHE_REQUIRE(rank >= 0 && rank < world, "shard_create: rank %d outside [0, %d)", rank, world);
    FHE_REQUIRE(max_batch >= 1, "shard_create: max_batch must be at least 1");
    FHE_REQUIRE(c->plan->bal, "shard_create: limb sharding needs the balanced two-pass NTT (2^13 <= N <= 2^16)");
    const uint32_t logw = host::ilog2((uint32_t)world);
    FHE_REQUIRE(logw + 8 <= c->logn, "shard_create: N/world = %u coefficients per rank; at least 256 are needed", c->n >> logw);
    const uint32_t L = c->L, R = c->R, K = c->K, A = L + R, W = L + K, dnum = c->dnum, alpha = c->alpha;
    FHE_REQUIRE((uint32_t)world <= W, "shard_create: more ranks (%d) than key limbs (%u)", world, W);
    DeviceGuard dev_guard(c->device);
    auto* s = new fhe_b200_shard();
    s->ctx = c; s->rank = rank; s->world = world; s->logw = (int)logw; s->max_batch = max_batch; s->nc = c->n >> logw;
    block_partition(A, world, s->a_begin);
    block_partition(W, world, s->w_begin);
    for (int r = 0; r < world; r++) {
        s->ca_max = std::max(s->ca_max, s->a_begin[r + 1] - s->a_begin[r]);
        s->cw_max = std::max(s->cw_max, s->w_begin[r + 1] - s->w_begin[r]);
    }
    const size_t N = c->n, Nc = s->nc, B = max_batch;
    size_t off = kCtlWords;
    s->off_ext = off; off += 4 * B * s->ca_max * N;
    s->off_d2 = off; off += 3 * B * A * Nc;
    s->off_dig = off; off += (size_t)dnum * B * s->cw_max * N;
    s->off_acc2 = off; off += 2 * B * W * Nc;
    s->slab_words = off;
    size_t wo = 0;
    s->off_d = wo; wo += 3 * B * s->ca_max * N;
    s->off_sr = wo; wo += 3 * B * R * Nc;
    s->off_sc = wo; wo += 3 * B * L * Nc;
    s->off_acc = wo; wo += 2 * B * s->cw_max * N;
    cudaError_t e = cudaMalloc(&s->slab, s->slab_words * 8);
    if (e == cudaSuccess) e = cudaMemset(s->slab, 0, kCtlWords * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s->ws, wo * 8);
    const size_t tab_words = 2 * ((size_t)R + L + (size_t)dnum * W);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_tabs, tab_words * 8);
    if (e != cudaSuccess) { set_error("shard_create: device allocation failed: %s", cudaGetErrorString(e)); fhe_b200_shard_destroy(s); return FHE_B200_ENOMEM; }
    if (const char* ev = getenv("FHE_B200_SHARD_TIMEOUT_MS")) { const long ms = atol(ev); if (ms > 0) s->timeout_ns = (unsigned long long)ms * 1000000ull; }
    (void)alpha;
    *out = s;
    return 0;
}

extern "C" int fhe_b200_shard_handle(const fhe_b200_shard* s, void* h_handle) {
    FHE_REQUIRE(s && h_handle, "shard_handle: null argument");
    DeviceGuard dev_guard(s->ctx->device);
    ShardHandle h;
    memset(&h, 0, sizeof(h));
    h.magic = kHandleMagic; h.pid = (int32_t)getpid(); h.device = s->ctx->device; h.world = (uint32_t)s->world;
    h.ptr = (uint64_t)(uintptr_t)s->slab; h.slab_words = s->slab_words;
    FHE_CUDA(cudaIpcGetMemHandle(&h.ipc, s->slab));
    memcpy(h_handle, &h, sizeof(h));
    return 0;
}

// h_handles: world handles of 128 bytes, in rank order (this rank's own included)
extern "C" int fhe_b200_shard_connect(fhe_b200_shard* s, const void* h_handles) {
    FHE_REQUIRE(s && h_handles, "shard_connect: null argument");
    FHE_REQUIRE(!s->connected, "shard_connect: already connected");
    fhe_b200_bfv* c = s->ctx;
    DeviceGuard dev_guard(c->device);
    const ShardHandle* hs = static_cast<const ShardHandle*>(h_handles);
    for (int r = 0; r < s->world; r++) {
        const ShardHandle& h = hs[r];
        FHE_REQUIRE(h.magic == kHandleMagic && h.world == (uint32_t)s->world && h.slab_words == s->slab_words,
                    "shard_connect: handle %d does not belong to this group (different parameters, batch or world size)", r);
        if (r == s->rank) { s->peer[r] = s->slab; continue; }
        if (h.pid == (int32_t)getpid()) {                        // same process: the pointer is valid here; make the device reachable
            if (h.device != c->device) {
                int can = 0;
                FHE_CUDA(cudaDeviceCanAccessPeer(&can, c->device, h.device));
                FHE_REQUIRE(can, "shard_connect: device %d cannot access device %d (no NVLink / P2P path)", c->device, h.device);
                const cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) FHE_CUDA(e);
                cudaGetLastError();
            }
            s->peer[r] = reinterpret_cast<uint64_t*>((uintptr_t)h.ptr);
        } else {
            void* p = nullptr;
            FHE_CUDA(cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess));
            s->peer[r] = static_cast<uint64_t*>(p); s->opened[r] = true;
        }
    }
    // scatter tables
    const uint32_t L = c->L, R = c->R, K = c->K, W = L + K, dnum = c->dnum, alpha = c->alpha;
    const size_t N = c->n, Nc = s->nc;
    auto owner = [&](const uint32_t* begin, uint32_t limb) { int r = 0; while (limb >= begin[r + 1]) r++; return r; };
    std::vector<uint64_t> tabs;
    uint64_t remote_words = 0;
    auto ext_entry = [&](uint32_t limb) {                        // plan limb `limb` of EXT on its owner
        const int h = owner(s->a_begin, limb);
        const uint32_t cnt = s->a_begin[h + 1] - s->a_begin[h];
        tabs.push_back((uint64_t)(uintptr_t)(s->peer[h] + s->off_ext + (size_t)(limb - s->a_begin[h]) * N + (size_t)s->rank * Nc));
        tabs.push_back((uint64_t)cnt * N);
        if (h != s->rank) remote_words += 4 * Nc;                // four input polynomials
    };
    const size_t o_q2r_out = tabs.size();
    for (uint32_t k = 0; k < R; k++) ext_entry(L + k);
    const size_t o_q2r_copy = tabs.size();
    for (uint32_t i = 0; i < L; i++) ext_entry(i);
    std::vector<size_t> o_up_out(dnum), o_up_copy(dnum);
    for (uint32_t dg = 0; dg < dnum; dg++) {
        auto dig_entry = [&](uint32_t limb) {                    // key limb `limb` of digit dg of DIG on its owner
            const int h = owner(s->w_begin, limb);
            const uint32_t cnt = s->w_begin[h + 1] - s->w_begin[h];
            tabs.push_back((uint64_t)(uintptr_t)(s->peer[h] + s->off_dig + ((size_t)dg * cnt + (limb - s->w_begin[h])) * N + (size_t)s->rank * Nc));
            tabs.push_back((uint64_t)dnum * cnt * N);
            if (h != s->rank) remote_words += Nc;
        };
        o_up_out[dg] = tabs.size();
        for (uint32_t i = 0; i < W; i++) if (i < dg * alpha || i >= (dg + 1) * alpha) dig_entry(i);
        o_up_copy[dg] = tabs.size();
        for (uint32_t i = dg * alpha; i < (dg + 1) * alpha; i++) dig_entry(i);
    }
    FHE_CUDA(cudaMemcpy(s->d_tabs, tabs.data(), tabs.size() * 8, cudaMemcpyHostToDevice));
    s->tab_q2r_out = s->d_tabs + o_q2r_out; s->tab_q2r_copy = s->d_tabs + o_q2r_copy;
    for (uint32_t dg = 0; dg < dnum; dg++) { s->tab_up_out.push_back(s->d_tabs + o_up_out[dg]); s->tab_up_copy.push_back(s->d_tabs + o_up_copy[dg]); }
    // the two scattered inverse transforms: all of this rank's limbs, every coefficient block but its own
    const uint32_t ca = s->a_begin[s->rank + 1] - s->a_begin[s->rank], cw = s->w_begin[s->rank + 1] - s->w_begin[s->rank];
    remote_words += (uint64_t)(3 * ca + 2 * cw) * (N - Nc);
    s->nvlink_words_per_op = remote_words;
    s->connected = true;
    return 0;
}

extern "C" int fhe_b200_shard_info(const fhe_b200_shard* s, uint32_t* coeff_begin, uint32_t* coeff_count, uint32_t* key_limb_begin,
                                   uint32_t* key_limb_count, uint32_t* ext_limb_begin, uint32_t* ext_limb_count, uint64_t* nvlink_bytes_per_op) {
    FHE_REQUIRE(s, "shard_info: null handle");
    if (coeff_begin) *coeff_begin = (uint32_t)s->rank * s->nc;
    if (coeff_count) *coeff_count = s->nc;
    if (key_limb_begin) *key_limb_begin = s->w_begin[s->rank];
    if (key_limb_count) *key_limb_count = s->w_begin[s->rank + 1] - s->w_begin[s->rank];
    if (ext_limb_begin) *ext_limb_begin = s->a_begin[s->rank];
    if (ext_limb_count) *ext_limb_count = s->a_begin[s->rank + 1] - s->a_begin[s->rank];
    if (nvlink_bytes_per_op) *nvlink_bytes_per_op = s->nvlink_words_per_op * 8;
    return 0;
}

// this rank's slice of a relinearisation / Galois key: [dnum][2][L+K][N] -> [dnum][2][cW][N]
extern "C" int fhe_b200_shard_slice_key(const fhe_b200_shard* s, const uint64_t* d_key, uint64_t* d_slice, void* stream) {
    FHE_REQUIRE(s && d_key && d_slice, "shard_slice_key: null argument");
    const fhe_b200_bfv* c = s->ctx;
    DeviceGuard dev_guard(c->device);
    const size_t N = c->n, W = c->L + c->K, cw = s->w_begin[s->rank + 1] - s->w_begin[s->rank];
    FHE_CUDA(cudaMemcpy2DAsync(d_slice, cw * N * 8, d_key + (size_t)s->w_begin[s->rank] * N, W * N * 8, cw * N * 8, 2 * (size_t)c->dnum,
                               cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

// synchronises the stream and reports a wait that gave up (a peer that never delivered)
extern "C" int fhe_b200_shard_check(fhe_b200_shard* s, void* stream) {
    FHE_REQUIRE(s, "shard_check: null handle");
    DeviceGuard dev_guard(s->ctx->device);
    ShardCtl h;
    FHE_CUDA(cudaMemcpyAsync(&h, s->slab, sizeof(ShardCtl), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    FHE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (h.err) {
        set_error("shard: rank %d timed out waiting for rank %u (operation %llu)", s->rank, h.err - 1, h.epoch);
        return FHE_B200_ESTATE;
    }
    return 0;
}

static int shard_signal(fhe_b200_shard* s, int phase, bool bump, cudaStream_t st) {
    PeerCtl pc;
    for (int r = 0; r < kMaxRanks; r++) pc.p[r] = r < s->world ? ctl_of(s->peer[r]) : nullptr;
    shard_signal_kernel<<<1, 32, 0, st>>>(ctl_of(s->slab), pc, phase, s->rank, s->world, bump ? 1 : 0);
    FHE_LAUNCH_CHECK();
    return 0;
}
static int shard_wait(fhe_b200_shard* s, int phase, cudaStream_t st) {
    shard_wait_kernel<<<1, 32, 0, st>>>(ctl_of(s->slab), phase, s->world, s->timeout_ns);
    FHE_LAUNCH_CHECK();
    return 0;
}

static inline uint32_t shard_grid(const fhe_b200_bfv* c, size_t items) {
    const size_t w = (items + 255) / 256, cap = (size_t)c->plan->sm_count * 16;
    return (uint32_t)(w < cap ? (w ? w : 1) : cap);
}

// The multiply as five stages; stage k begins with the wait for phase k-1 and ends with the signal of phase k:
//   0: Q -> R extension (scattered by limb)                          | signal 0
//   1: wait 0 | NTT, tensor product, inverse NTT (scattered by coefficient block) | signal 1
//   2: wait 1 | scale-and-round, R -> Q, ModUp (scattered by limb)               | signal 2
//   3: wait 2 | NTT, key inner product, inverse NTT (scattered by coefficient block) | signal 3
//   4: wait 3 | ModDown + add
// d_a, d_b, d_out: sharded ciphertexts [batch][2][L][Nc]; d_key: this rank's key slice [dnum][2][cW][N]
static int sharded_stage(fhe_b200_shard* s, int stage, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_key, uint64_t* d_out,
                         uint32_t batch, cudaStream_t st) {
    fhe_b200_bfv* c = s->ctx;
    const uint32_t L = c->L, R = c->R, A = L + R, W = L + c->K, dnum = c->dnum, B = batch, nc = s->nc;
    const size_t N = c->n, Nc = nc;
    const uint32_t ab = s->a_begin[s->rank], ca = s->a_begin[s->rank + 1] - ab, wb = s->w_begin[s->rank], cw = s->w_begin[s->rank + 1] - wb;
    const LimbParams* prm = c->plan->d_params;
    const bool square = d_a == d_b;
    uint64_t* ext = s->slab + s->off_ext; uint64_t* d2 = s->slab + s->off_d2; uint64_t* dig = s->slab + s->off_dig; uint64_t* acc2 = s->slab + s->off_acc2;
    uint64_t* dl = s->ws + s->off_d; uint64_t* sR = s->ws + s->off_sr; uint64_t* sc = s->ws + s->off_sc; uint64_t* accl = s->ws + s->off_acc;
    const bool fused = getenv("FHE_B200_FUSED_TILE") && atoi(getenv("FHE_B200_FUSED_TILE")) != 0 && fused_tile_supported(c->plan, dnum);   // opt-in, see bfv.cu
    BalScatter bs; memset(&bs, 0, sizeof(bs));
    bs.log_blocks = (uint32_t)s->logw; bs.nc = nc;
    switch (stage) {
    case 0:
        // exact extension Q -> Q u R of this rank's coefficients; every limb goes to its owner (Q limbs passed through)
        for (int o = 0; o < (square ? 1 : 2); o++) {
            LcView v; v.in = o ? d_b : d_a; v.in_stride = (size_t)L * Nc;
            v.out_tab = s->tab_q2r_out; v.out_poly0 = (size_t)o * 2 * B;
            v.copy_out = ext; v.copy_tab = s->tab_q2r_copy; v.copy_poly0 = (size_t)o * 2 * B;
            FHE_TRY(lincomb_launch(c->q2r, v, nc, 2 * B, st));
        }
        return shard_signal(s, 0, true, st);
    case 1: {
        FHE_TRY(shard_wait(s, 0, st));
        // own limbs: NTT, tensor product, inverse NTT whose last pass hands every coefficient block to its owner
        const uint32_t planes = square ? 2 : 4;
        for (int r = 0; r < s->world; r++) bs.base[r] = (uint64_t)(uintptr_t)(s->peer[r] + s->off_d2);
        bs.limbs_total = A; bs.limb_off = ab;
        if (fused) {
            // column pass, then one kernel for tile pass x4 + tensor product + inverse tile pass x3, then the scattering column pass
            FHE_TRY(launch_ntt_pass_a(c->plan, ext, ext, planes * B, ab, ca, false, st));
            FusedTile t; t.in = ext; t.out = dl; t.limb_begin = ab; t.limb_count = ca; t.nb = B; t.square = square;
            const size_t cn = (size_t)ca * N;
            t.in_plane[0] = 0; t.in_plane[1] = cn; t.in_plane[2] = 2 * (size_t)B * cn; t.in_plane[3] = (2 * (size_t)B + 1) * cn; t.in_poly = 2 * cn;
            t.out_plane[0] = 0; t.out_plane[1] = cn; t.out_plane[2] = 2 * cn; t.out_poly = 3 * cn;
            FHE_TRY(launch_fused_tile(c->plan, 0, t, st));
            FHE_TRY(launch_ntt_pass_a(c->plan, dl, dl, 3 * B, ab, ca, true, st, &bs));
        } else {
            FHE_TRY(launch_ntt(c->plan, ext, ext, planes * B, ab, ca, false, st));
            const size_t per = (size_t)B * ca * N / 2;
            shard_tensor_kernel<<<shard_grid(c, per), 256, 0, st>>>((ulonglong2*)dl, (const ulonglong2*)ext, prm, c->logn, ab, ca, B,
                                                                    square ? 0 : (size_t)B * 2 * ca * N / 2);
            FHE_LAUNCH_CHECK();
            FHE_TRY(launch_ntt_inverse_scatter(c->plan, dl, 3 * B, ab, ca, bs, st));
        }
        return shard_signal(s, 1, false, st);
    }
    case 2:
        FHE_TRY(shard_wait(s, 1, st));
        // own coefficients: round(t/Q .) in basis R, exact conversion R -> Q, ModUp of d2 with every key limb going to its owner
        { LcView v; v.in = d2; v.in_stride = (size_t)A * Nc; v.extra = d2 + (size_t)L * Nc; v.extra_stride = (size_t)A * Nc; v.out = sR; v.out_stride = (size_t)R * Nc;
          FHE_TRY(lincomb_launch(c->scale, v, nc, 3 * B, st)); }
        { LcView v; v.in = sR; v.in_stride = (size_t)R * Nc; v.out = sc; v.out_stride = (size_t)L * Nc;
          FHE_TRY(lincomb_launch(c->r2q, v, nc, 3 * B, st)); }
        for (uint32_t dg = 0; dg < dnum; dg++) {
            LcView v; v.in = sc + 2 * (size_t)L * Nc; v.in_stride = 3 * (size_t)L * Nc; v.src_idx = c->d_idx_grp[dg];
            v.out_tab = s->tab_up_out[dg]; v.copy_out = dig; v.copy_tab = s->tab_up_copy[dg];
            FHE_TRY(lincomb_launch(c->modup[dg], v, nc, B, st));
        }
        return shard_signal(s, 2, false, st);
    case 3:
        FHE_TRY(shard_wait(s, 2, st));
        // own key limbs: NTT of the digits, inner product with the key slice, inverse NTT with the scattering last pass
        for (int r = 0; r < s->world; r++) bs.base[r] = (uint64_t)(uintptr_t)(s->peer[r] + s->off_acc2);
        bs.limbs_total = W; bs.limb_off = wb;
        if (fused) {
            FHE_TRY(launch_ntt_pass_a(c->plan, dig, dig, dnum * B, wb, cw, false, st));
            FusedTile t; t.in = dig; t.out = accl; t.limb_begin = wb; t.limb_count = cw; t.nb = B; t.dnum = dnum;
            const size_t cn = (size_t)cw * N;
            for (uint32_t d = 0; d < dnum; d++) t.in_plane[d] = (size_t)d * cn;
            t.in_poly = (size_t)dnum * cn; t.out_plane[0] = 0; t.out_plane[1] = cn; t.out_poly = 2 * cn;
            t.key = d_key; t.key_poly = cn;
            FHE_TRY(launch_fused_tile(c->plan, 1, t, st));
            FHE_TRY(launch_ntt_pass_a(c->plan, accl, accl, 2 * B, wb, cw, true, st, &bs));
        } else {
            FHE_TRY(launch_ntt(c->plan, dig, dig, dnum * B, wb, cw, false, st));
            const size_t per = (size_t)B * cw * N / 2;
            shard_ks_inner_kernel<<<shard_grid(c, per), 256, 0, st>>>((ulonglong2*)accl, (const ulonglong2*)dig, (const ulonglong2*)d_key, prm, c->logn,
                                                                      wb, cw, dnum, B);
            FHE_LAUNCH_CHECK();
            FHE_TRY(launch_ntt_inverse_scatter(c->plan, accl, 2 * B, wb, cw, bs, st));
        }
        return shard_signal(s, 3, false, st);
    case 4:
        FHE_TRY(shard_wait(s, 3, st));
        // own coefficients: ModDown, added onto (d0, d1)
        for (int p = 0; p < 2; p++) {
            const uint64_t* a2 = acc2 + (size_t)p * W * Nc;
            LcView v; v.in = a2; v.in_stride = 2 * (size_t)W * Nc; v.src_idx = c->d_idx_p;
            v.sub = a2; v.sub_stride = 2 * (size_t)W * Nc; v.epi_scalar = c->d_pinv; v.epi_scalar_shoup = c->d_pinv_s;
            v.add = sc + (size_t)p * L * Nc; v.add_stride = 3 * (size_t)L * Nc;
            v.out = d_out + (size_t)p * L * Nc; v.out_stride = 2 * (size_t)L * Nc;
            FHE_TRY(lincomb_launch(c->moddown, v, nc, B, st));
        }
        FHE_CUDA(cudaGetLastError());
        return 0;
    }
    set_error("bfv_multiply_relin_sharded: stage %d outside [0, 4]", stage);
    return FHE_B200_EINVAL;
}

static int sharded_check_args(fhe_b200_shard* s, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_key, uint64_t* d_out, uint32_t batch) {
    FHE_REQUIRE(s && d_a && d_b && d_key && d_out, "bfv_multiply_relin_sharded: null argument");
    FHE_REQUIRE(s->connected, "bfv_multiply_relin_sharded: the shard group is not connected (fhe_b200_shard_connect)");
    FHE_REQUIRE(batch <= s->max_batch, "bfv_multiply_relin_sharded: batch %u exceeds the group's max_batch %u", batch, s->max_batch);
    return 0;
}

extern "C" int fhe_b200_bfv_multiply_relin_sharded(fhe_b200_shard* s, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_key,
                                                   uint64_t* d_out, uint32_t batch, void* stream) {
    FHE_TRY(sharded_check_args(s, d_a, d_b, d_key, d_out, batch));
    if (!batch) return 0;
    DeviceGuard dev_guard(s->ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    // The second call with the same arguments captures the five stages into a CUDA graph; later calls replay it (FHE_B200_SHARD_GRAPH=0:
    // always launch kernel by kernel).  The first call runs eagerly, so one-off work (function attributes, allocations) is never captured.
    const char* ge = getenv("FHE_B200_SHARD_GRAPH");
    const bool want_graph = !(ge && atoi(ge) == 0) && st != nullptr;
    const char* fe = getenv("FHE_B200_FUSED_TILE");
    const fhe_b200_shard::GraphKey key{d_a, d_b, d_key, d_out, batch, (fe && atoi(fe) != 0) ? 1 : 0};
    fhe_b200_shard::GraphEntry* ent = nullptr;
    if (want_graph) {
        for (auto& g : s->graphs)
            if (g.k.a == key.a && g.k.b == key.b && g.k.key == key.key && g.k.out == key.out && g.k.batch == key.batch && g.k.fused == key.fused) { ent = &g; break; }
        if (!ent) {
            if (s->graphs.size() >= 16) { for (auto& g : s->graphs) if (g.exec) cudaGraphExecDestroy(g.exec); s->graphs.clear(); }
            s->graphs.push_back({key, nullptr, 0});
            ent = &s->graphs.back();
        }
        if (ent->exec) { FHE_CUDA(cudaGraphLaunch(ent->exec, st)); count_launch(); return 0; }
        if (++ent->seen == 2) {
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                int rc = 0;
                for (int stage = 0; stage < 5 && !rc; stage++) rc = sharded_stage(s, stage, d_a, d_b, d_key, d_out, batch, st);
                const cudaError_t ce = cudaStreamEndCapture(st, &graph);
                if (!rc && ce == cudaSuccess && graph && cudaGraphInstantiate(&ent->exec, graph, 0) == cudaSuccess) {
                    cudaGraphDestroy(graph);
                    FHE_CUDA(cudaGraphLaunch(ent->exec, st)); count_launch();
                    return 0;
                }
                if (graph) cudaGraphDestroy(graph);
                ent->exec = nullptr; ent->seen = 1000;              // capture not possible here: stay with plain launches for this key
                cudaGetLastError();
                if (rc) return rc;
            } else cudaGetLastError();
        }
    }
    for (int stage = 0; stage < 5; stage++) FHE_TRY(sharded_stage(s, stage, d_a, d_b, d_key, d_out, batch, st));
    return 0;
}

// One stage of the multiply (0..4, see sharded_stage).  For a host that drives SEVERAL ranks on ONE device (the single-GPU tests):
// kernels that wait for one another must not be separate launches on one GPU -- nothing guarantees that they run at the same time --
// so such a host issues stage k for every rank before stage k+1 for any rank, and every wait finds its flags already set.
extern "C" int fhe_b200_bfv_multiply_relin_sharded_stage(fhe_b200_shard* s, int stage, const uint64_t* d_a, const uint64_t* d_b,
                                                         const uint64_t* d_key, uint64_t* d_out, uint32_t batch, void* stream) {
    FHE_TRY(sharded_check_args(s, d_a, d_b, d_key, d_out, batch));
    if (!batch) return 0;
    DeviceGuard dev_guard(s->ctx->device);
    return sharded_stage(s, stage, d_a, d_b, d_key, d_out, batch, (cudaStream_t)stream);
}
