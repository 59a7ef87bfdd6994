// Multi-GPU exchanges (one process per GPU) over NVLink 5 / NVSwitch.
//
// The probe and the block-level build shard by file with no exchange; the two real exchange steps of
// the path are:
//   bsg_or_reduce[_device]       bitwise OR of equal-(m,k) partial file-level bitsets built from disjoint
//                                shards of one file's entries (flush.go:221,253 builds that filter from the
//                                union of the entry sets; OR of partials is the same bitset).
//   bsg_allgather_masks[_device] all-gather of the per-rank candidate masks (query_exec.go:572-615 gives
//                                one bit per block; blocks are sharded by file).
//
// Two implementations, chosen once at bsg_comm_init:
//   peer memory (default on one NVSwitch node): every buffer that takes part is SYMMETRIC (bsg_comm_alloc:
//       same size on every rank, every rank maps every peer's copy through CUDA IPC).  Each collective
//       is ONE kernel: ranks announce "input ready" with a flag store into every peer, wait for all
//       peers, then rank r ORs slice r of all W partials straight out of peer memory (NVLink loads) and
//       stores the result into slice r of every peer (NVLink stores); a second flag round closes it.
//       Traffic per rank = 2*(W-1)/W * bytes, the reduce-scatter + all-gather bound, with no staging
//       copy, no per-call allocation and no host round trip.  NCCL has no OR reduction at all.
//   NCCL (fallback when IPC mapping is unavailable): all-to-all of 1/W slices (grouped ncclSend/ncclRecv)
//       into a cached receive buffer, ONE OR kernel over the W received slices, ncclAllGather.
// NCCL is dlopen'ed so single-GPU users carry no NCCL dependency; it also carries the bootstrap
// (unique id rendezvous, exchange of the IPC handles).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "bsg_internal.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;
std::string g_nccl_err;

void load_nccl() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) { g_nccl_err = dlerror() ? dlerror() : "libnccl not found"; return; }
#define LOAD(field, sym)                                                            \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.handle, sym)); \
    if (!g_nccl.field) { g_nccl_err = std::string("missing symbol ") + sym; g_nccl.handle = nullptr; return; }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(AllGather, "ncclAllGather")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
}

constexpr int kMaxPeers = 8;             // one NVSwitch node
constexpr size_t kSymAlign = 2u << 20;   // symmetric buffers are whole 2 MiB blocks (one IPC handle each)

// Symmetric control block: flags the peers write into (system-scope stores over NVLink).
struct Ctl {
    uint32_t ready[kMaxPeers];  // ready[p] = last collective for which rank p announced its input / free output
    uint32_t done[kMaxPeers];   // done[p]  = last collective whose writes from rank p have landed here
    uint32_t counter;           // CTA completion counter of this rank's collective kernel
    uint32_t selftest;
    uint32_t pad[14];
};

struct SymBuf {
    uint8_t* local = nullptr;
    size_t bytes = 0;
    uint8_t* peer[kMaxPeers] = {};  // peer[rank] == local
};

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    int device = 0;
    cudaStream_t stream = nullptr;  // setup + host-pointer wrappers
    bool p2p = false;
    SymBuf ctl;                     // the control blocks
    uint32_t epoch = 0;
    std::vector<SymBuf> bufs;
    uint64_t *d_send = nullptr, *d_recv = nullptr;  // NCCL path scratch (cached, grows)
    size_t cap_scratch = 0;
    uint8_t* stage = nullptr;       // symmetric staging for the host-pointer wrappers (cached, grows)
    size_t cap_stage = 0;
    uint8_t* d_xchg = nullptr;      // 64 B x world: IPC handle exchange
    std::mutex mu;
    uint64_t nvlink_bytes = 0;      // bytes this rank moved over NVLink in the last collective (reporting)
    uint64_t timeout_ns = 120ull * 1000000000ull;  // BSG_COMM_TIMEOUT_S: longest a collective kernel waits for a peer
};

}  // namespace

// api.cu owns bsg_ctx; these accessors keep this file free of its layout.
extern "C" int bsg_ctx_device_internal(bsg_ctx* ctx);
extern "C" void** bsg_ctx_comm_slot_internal(bsg_ctx* ctx);
extern "C" void* bsg_ctx_stream_internal(bsg_ctx* ctx);
extern "C" int bsg_set_last_error_internal(int code, const char* msg);

#define NCCL_TRY(expr)                                                                         \
    do {                                                                                       \
        ncclResult_t _r = (expr);                                                              \
        if (_r != ncclSuccess)                                                                 \
            return bsg_set_last_error_internal(BSG_ERR_COMM, (std::string(#expr) + ": " + g_nccl.GetErrorString(_r)).c_str()); \
    } while (0)
#define CU_TRY(expr)                                                                           \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return bsg_set_last_error_internal(BSG_ERR_CUDA, (std::string(#expr) + ": " + cudaGetErrorString(_e)).c_str()); \
    } while (0)

static int need_nccl() {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.handle) return bsg_set_last_error_internal(BSG_ERR_UNSUPPORTED, ("NCCL unavailable: " + g_nccl_err).c_str());
    return BSG_OK;
}

// ------------------------------------------------------------------ kernels ---
namespace {

struct PeerPtrs { uint8_t* p[kMaxPeers]; };
struct Sig {
    Ctl* mine;                 // this rank's control block
    Ctl* peer[kMaxPeers];      // every rank's control block as mapped here (peer[rank] == mine)
    uint64_t timeout_ns;       // a flag wait longer than this traps the kernel (0 = wait for ever)
    uint32_t epoch;
    int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Peer loads must not come from a stale L1 line and need no L1 allocation.
__device__ __forceinline__ uint4 ld_peer_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ uint64_t comm_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Spins until *flag has reached `epoch`.  A peer that never arrives (its process died, its stream is stuck)
// must not leave this GPU spinning for ever: after timeout_ns the kernel traps, the context reports a launch
// failure at the next CUDA call and the caller fails loudly instead of hanging (BSG_COMM_TIMEOUT_S, default 120).
__device__ __forceinline__ void comm_wait_flag(const uint32_t* flag, uint32_t epoch, uint64_t timeout_ns) {
    if (static_cast<int32_t>(ld_acquire_sys(flag) - epoch) >= 0) return;
    const uint64_t t0 = comm_timer_ns();
    uint32_t polls = 0;
    while (static_cast<int32_t>(ld_acquire_sys(flag) - epoch) < 0) {
        if ((++polls & 0xfffu) == 0 && timeout_ns && comm_timer_ns() - t0 > timeout_ns) asm volatile("trap;");
    }
}

// "my input is ready / my output may be overwritten": announce to every peer, then wait for all of them.
// Executed by every CTA (the announcement by CTA 0 only); ends with a CTA barrier.
__device__ __forceinline__ void comm_ready_barrier(const Sig& s) {
    if (blockIdx.x == 0 && threadIdx.x < static_cast<uint32_t>(s.world)) {
        __threadfence_system();
        st_release_sys(&s.peer[threadIdx.x]->ready[s.rank], s.epoch);
    }
    if (threadIdx.x < static_cast<uint32_t>(s.world)) comm_wait_flag(&s.mine->ready[threadIdx.x], s.epoch, s.timeout_ns);
    __syncthreads();
}

// After this CTA's stores into peer memory: the last CTA of the grid tells every peer "my writes have
// landed" and waits until every peer has said the same, so when the kernel ends the result is complete here.
__device__ __forceinline__ void comm_done_barrier(const Sig& s) {
    __threadfence_system();
    __syncthreads();
    __shared__ uint32_t s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(&s.mine->counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) s.mine->counter = 0;
    if (threadIdx.x < static_cast<uint32_t>(s.world)) {
        __threadfence_system();
        st_release_sys(&s.peer[threadIdx.x]->done[s.rank], s.epoch);
        comm_wait_flag(&s.mine->done[threadIdx.x], s.epoch, s.timeout_ns);
    }
    __syncthreads();
}

__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// In-place OR across ranks.  bufs.p[r] = rank r's partial (n_words uint64, 16-byte aligned).  Rank `rank`
// owns the 16-byte units [lo16, hi16): it ORs them out of every peer and writes the result to every peer.
__global__ void __launch_bounds__(512)
or_reduce_p2p_kernel(PeerPtrs bufs, uint64_t n_words, Sig sig) {
    comm_ready_barrier(sig);
    const int W = sig.world;
    const uint64_t total16 = n_words >> 1;
    const uint64_t per = (total16 + W - 1) / W;
    const uint64_t lo = umin64(total16, per * sig.rank), hi = umin64(total16, lo + per);
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = lo + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < hi; i += stride) {
        uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p) {
            if (p < W) {
                const uint4 v = ld_peer_u4(reinterpret_cast<const uint4*>(bufs.p[p]) + i);
                acc.x |= v.x; acc.y |= v.y; acc.z |= v.z; acc.w |= v.w;
            }
        }
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p)
            if (p < W) reinterpret_cast<uint4*>(bufs.p[p])[i] = acc;
    }
    if ((n_words & 1) && sig.rank == W - 1 && blockIdx.x == 0 && threadIdx.x == 0) {  // odd tail word
        uint64_t acc = 0;
        for (int p = 0; p < W; ++p) acc |= *reinterpret_cast<volatile uint64_t*>(bufs.p[p] + (n_words - 1) * 8);
        for (int p = 0; p < W; ++p) *reinterpret_cast<volatile uint64_t*>(bufs.p[p] + (n_words - 1) * 8) = acc;
    }
    comm_done_barrier(sig);
}

// All-gather: every rank stores its n_words into slot `rank` of every peer's output.
__global__ void __launch_bounds__(256)
allgather_p2p_kernel(const uint64_t* __restrict__ local, uint64_t n_words, PeerPtrs all, Sig sig) {
    comm_ready_barrier(sig);
    const int W = sig.world;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_words; i += stride) {
        const uint64_t v = local[i];
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p)
            if (p < W) reinterpret_cast<uint64_t*>(all.p[p])[static_cast<uint64_t>(sig.rank) * n_words + i] = v;
    }
    comm_done_barrier(sig);
}

// NCCL path: dst[i] = OR over the W received slices.
__global__ void __launch_bounds__(256)
or_slices_kernel(uint64_t* __restrict__ dst, const uint64_t* __restrict__ recv, uint64_t sl, int W) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < sl; i += stride) {
        uint64_t acc = 0;
        for (int p = 0; p < W; ++p) acc |= recv[static_cast<uint64_t>(p) * sl + i];
        dst[i] = acc;
    }
}

}  // namespace

// ------------------------------------------------------------------- set-up ---
extern "C" int bsg_comm_unique_id(uint8_t out_id[128]) {
    if (!out_id) return bsg_set_last_error_internal(BSG_ERR_INVALID, "out_id is NULL");
    int rc = need_nccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(out_id, &id, 128);
    return BSG_OK;
}

static void comm_release(Comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    auto close_buf = [&](SymBuf& b) {
        for (int p = 0; p < c->world && p < kMaxPeers; ++p)
            if (p != c->rank && b.peer[p]) cudaIpcCloseMemHandle(b.peer[p]);
        if (b.local) cudaFree(b.local);
        b = SymBuf();
    };
    for (SymBuf& b : c->bufs) close_buf(b);
    close_buf(c->ctl);
    cudaFree(c->d_send);
    cudaFree(c->d_recv);
    cudaFree(c->d_xchg);
    if (c->comm && g_nccl.handle) g_nccl.CommDestroy(c->comm);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// min over ranks of `ok` (collective): every rank must take the same decision.
static int agree(Comm* c, int ok, int* all_ok) {
    int* d = reinterpret_cast<int*>(c->d_xchg);
    CU_TRY(cudaMemcpyAsync(d, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(g_nccl.AllReduce(d, d, 1, ncclInt32, ncclMin, c->comm, c->stream));
    CU_TRY(cudaMemcpyAsync(all_ok, d, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return BSG_OK;
}

// Allocates `bytes` (rounded up to 2 MiB) on this rank and maps every peer's allocation (collective).
// *mapped = 0 when some rank could not map a peer (the caller decides whether that is fatal).
static int sym_alloc(Comm* c, size_t bytes, SymBuf* out, int* mapped) {
    *out = SymBuf();
    out->bytes = (std::max<size_t>(bytes, 1) + kSymAlign - 1) / kSymAlign * kSymAlign;
    CU_TRY(cudaMalloc(reinterpret_cast<void**>(&out->local), out->bytes));
    CU_TRY(cudaMemsetAsync(out->local, 0, out->bytes, c->stream));
    out->peer[c->rank] = out->local;
    int ok = 1;
    cudaIpcMemHandle_t mine;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    if (cudaIpcGetMemHandle(&mine, out->local) != cudaSuccess) { ok = 0; memset(&mine, 0, sizeof(mine)); cudaGetLastError(); }
    std::vector<cudaIpcMemHandle_t> all(c->world);
    CU_TRY(cudaMemcpyAsync(c->d_xchg + 64 * c->rank, &mine, 64, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(g_nccl.AllGather(c->d_xchg + 64 * c->rank, c->d_xchg, 64, ncclUint8, c->comm, c->stream));
    CU_TRY(cudaMemcpyAsync(all.data(), c->d_xchg, 64 * c->world, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    for (int p = 0; p < c->world && ok; ++p) {
        if (p == c->rank) continue;
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
        out->peer[p] = static_cast<uint8_t*>(ptr);
    }
    int rc = agree(c, ok, mapped);
    if (rc) return rc;
    if (!*mapped)
        for (int p = 0; p < c->world; ++p)
            if (p != c->rank && out->peer[p]) { cudaIpcCloseMemHandle(out->peer[p]); out->peer[p] = nullptr; }
    return BSG_OK;
}

extern "C" int bsg_comm_init(bsg_ctx* ctx, int rank, int world, const uint8_t nccl_unique_id[128]) {
    if (!ctx || !nccl_unique_id || world < 1 || rank < 0 || rank >= world)
        return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_comm_init: bad argument");
    int rc = need_nccl();
    if (rc) return rc;
    void** slot = bsg_ctx_comm_slot_internal(ctx);
    if (*slot) return bsg_set_last_error_internal(BSG_ERR_INVALID, "communicator already initialised");
    Comm* c = new Comm();
    c->rank = rank;
    c->world = world;
    if (const char* t = getenv("BSG_COMM_TIMEOUT_S")) c->timeout_ns = static_cast<uint64_t>(std::max(0.0, atof(t)) * 1e9);
    c->device = bsg_ctx_device_internal(ctx);
    auto body = [&]() -> int {
        CU_TRY(cudaSetDevice(c->device));
        CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        ncclUniqueId id;
        memcpy(&id, nccl_unique_id, 128);
        NCCL_TRY(g_nccl.CommInitRank(&c->comm, world, id, rank));
        CU_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_xchg), 64 * static_cast<size_t>(std::max(world, 2))));
        const char* env = getenv("BSG_COMM_P2P");
        const bool want_p2p = world <= kMaxPeers && !(env && atoi(env) == 0);
        if (!want_p2p) return BSG_OK;
        // control blocks + a self test of the mapping: every rank reads what each peer wrote into its own block
        int mapped = 0;
        int r = sym_alloc(c, sizeof(Ctl), &c->ctl, &mapped);
        if (r) return r;
        int ok = mapped;
        if (mapped) {
            const uint32_t tag = 0xB100u + rank;
            CU_TRY(cudaMemcpyAsync(c->ctl.local + offsetof(Ctl, selftest), &tag, 4, cudaMemcpyHostToDevice, c->stream));
            int dummy = 0;
            r = agree(c, 1, &dummy);  // doubles as a barrier: every rank has written its tag
            if (r) return r;
            for (int p = 0; p < world && ok; ++p) {
                uint32_t got = 0;
                if (cudaMemcpyAsync(&got, c->ctl.peer[p] + offsetof(Ctl, selftest), 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                    cudaStreamSynchronize(c->stream) != cudaSuccess || got != 0xB100u + p) { ok = 0; cudaGetLastError(); }
            }
        }
        int all_ok = 0;
        r = agree(c, ok, &all_ok);
        if (r) return r;
        c->p2p = all_ok != 0;
        return BSG_OK;
    };
    rc = body();
    if (rc) { comm_release(c); return rc; }
    *slot = c;
    return BSG_OK;
}

extern "C" void bsg_comm_destroy_internal(void* comm) { comm_release(static_cast<Comm*>(comm)); }

static Comm* comm_of(bsg_ctx* ctx) { return ctx ? static_cast<Comm*>(*bsg_ctx_comm_slot_internal(ctx)) : nullptr; }

extern "C" int bsg_comm_info(bsg_ctx* ctx, int* rank, int* world, int* peer_memory, uint64_t* last_nvlink_bytes) {
    Comm* c = comm_of(ctx);
    if (!c) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_comm_init has not been called");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (peer_memory) *peer_memory = c->p2p ? 1 : 0;
    if (last_nvlink_bytes) *last_nvlink_bytes = c->nvlink_bytes;
    return BSG_OK;
}

extern "C" int bsg_comm_alloc(bsg_ctx* ctx, size_t bytes, void** out_dev) {
    Comm* c = comm_of(ctx);
    if (!c || !out_dev) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_comm_alloc: bad argument / no communicator");
    *out_dev = nullptr;
    std::lock_guard<std::mutex> lk(c->mu);
    CU_TRY(cudaSetDevice(c->device));
    SymBuf b;
    if (c->p2p) {
        int mapped = 0;
        int rc = sym_alloc(c, bytes, &b, &mapped);
        if (rc) { if (b.local) cudaFree(b.local); return rc; }
        if (!mapped) { cudaFree(b.local); return bsg_set_last_error_internal(BSG_ERR_COMM, "a rank could not map a peer's buffer (CUDA IPC)"); }
    } else {
        b.bytes = (std::max<size_t>(bytes, 1) + kSymAlign - 1) / kSymAlign * kSymAlign;
        CU_TRY(cudaMalloc(reinterpret_cast<void**>(&b.local), b.bytes));
        CU_TRY(cudaMemset(b.local, 0, b.bytes));
        b.peer[c->rank] = b.local;
    }
    c->bufs.push_back(b);
    *out_dev = b.local;
    return BSG_OK;
}

extern "C" int bsg_comm_free(bsg_ctx* ctx, void* dev) {
    Comm* c = comm_of(ctx);
    if (!c) return bsg_set_last_error_internal(BSG_ERR_INVALID, "no communicator");
    if (!dev) return BSG_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CU_TRY(cudaSetDevice(c->device));
    for (size_t i = 0; i < c->bufs.size(); ++i) {
        if (c->bufs[i].local != dev) continue;
        CU_TRY(cudaDeviceSynchronize());
        for (int p = 0; p < c->world && p < kMaxPeers; ++p)
            if (p != c->rank && c->bufs[i].peer[p]) cudaIpcCloseMemHandle(c->bufs[i].peer[p]);
        cudaFree(c->bufs[i].local);
        c->bufs.erase(c->bufs.begin() + i);
        return BSG_OK;
    }
    return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_comm_free: not a bsg_comm_alloc buffer");
}

// Locates [ptr, ptr+bytes) inside a symmetric buffer and returns the same range on every peer.
static bool peers_of(Comm* c, const void* ptr, size_t bytes, PeerPtrs* out) {
    const uint8_t* p8 = static_cast<const uint8_t*>(ptr);
    for (const SymBuf& b : c->bufs) {
        if (p8 < b.local || p8 + bytes > b.local + b.bytes) continue;
        const size_t off = p8 - b.local;
        for (int p = 0; p < kMaxPeers; ++p) out->p[p] = (p < c->world && b.peer[p]) ? b.peer[p] + off : nullptr;
        return true;
    }
    return false;
}

static Sig make_sig(Comm* c) {
    Sig s;
    s.mine = reinterpret_cast<Ctl*>(c->ctl.local);
    for (int p = 0; p < kMaxPeers; ++p) s.peer[p] = p < c->world ? reinterpret_cast<Ctl*>(c->ctl.peer[p]) : nullptr;
    s.epoch = ++c->epoch;
    s.timeout_ns = c->timeout_ns;
    s.rank = c->rank;
    s.world = c->world;
    return s;
}

static int ensure_scratch(Comm* c, size_t words) {
    if (words <= c->cap_scratch) return BSG_OK;
    cudaFree(c->d_send);
    cudaFree(c->d_recv);
    c->d_send = c->d_recv = nullptr;
    c->cap_scratch = 0;
    CU_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_send), words * 8));
    CU_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_recv), words * 8));
    c->cap_scratch = words;
    return BSG_OK;
}

// ---------------------------------------------------------------- OR-reduce ---
static int or_reduce_on(Comm* c, uint64_t* d_words, uint64_t n_words, cudaStream_t s) {
    const uint64_t W = static_cast<uint64_t>(c->world);
    c->nvlink_bytes = 2 * (W - 1) * ((n_words + W - 1) / W) * 8;
    if (W == 1) return BSG_OK;
    if (c->p2p) {
        PeerPtrs bufs;
        if ((reinterpret_cast<uintptr_t>(d_words) & 15) || !peers_of(c, d_words, n_words * 8, &bufs))
            return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_or_reduce_device: d_words must be 16-byte aligned memory from bsg_comm_alloc");
        const Sig sig = make_sig(c);
        const uint64_t per = ((n_words >> 1) + W - 1) / W;
        int blocks = static_cast<int>(std::min<uint64_t>(148 * 2, (per + 511) / 512));
        if (blocks < 1) blocks = 1;
        or_reduce_p2p_kernel<<<blocks, 512, 0, s>>>(bufs, n_words, sig);
        CU_TRY(cudaGetLastError());
        return BSG_OK;
    }
    // NCCL: all-to-all of 1/W slices, one OR kernel over the W received slices, all-gather
    const uint64_t sl = (n_words + W - 1) / W;
    int rc = ensure_scratch(c, sl * W);
    if (rc) return rc;
    CU_TRY(cudaMemsetAsync(c->d_send, 0, sl * W * 8, s));
    CU_TRY(cudaMemcpyAsync(c->d_send, d_words, n_words * 8, cudaMemcpyDeviceToDevice, s));
    NCCL_TRY(g_nccl.GroupStart());
    for (uint64_t p = 0; p < W; ++p) {
        NCCL_TRY(g_nccl.Send(c->d_send + p * sl, sl, ncclUint64, static_cast<int>(p), c->comm, s));
        NCCL_TRY(g_nccl.Recv(c->d_recv + p * sl, sl, ncclUint64, static_cast<int>(p), c->comm, s));
    }
    NCCL_TRY(g_nccl.GroupEnd());
    uint64_t* mine = c->d_send + static_cast<uint64_t>(c->rank) * sl;
    or_slices_kernel<<<static_cast<int>(std::min<uint64_t>(148 * 4, (sl + 255) / 256)), 256, 0, s>>>(mine, c->d_recv, sl, c->world);
    CU_TRY(cudaGetLastError());
    NCCL_TRY(g_nccl.AllGather(mine, c->d_recv, sl, ncclUint64, c->comm, s));
    CU_TRY(cudaMemcpyAsync(d_words, c->d_recv, n_words * 8, cudaMemcpyDeviceToDevice, s));
    return BSG_OK;
}

extern "C" int bsg_or_reduce_device(bsg_ctx* ctx, uint64_t* d_words, uint64_t n_words) {
    Comm* c = comm_of(ctx);
    if (!c || (n_words && !d_words)) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_or_reduce_device: bad argument / no communicator");
    if (n_words == 0) return BSG_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CU_TRY(cudaSetDevice(c->device));
    return or_reduce_on(c, d_words, n_words, static_cast<cudaStream_t>(bsg_ctx_stream_internal(ctx)));
}

// Grows the symmetric staging area of the host-pointer wrappers (collective: every rank passes the same size).
static int ensure_stage(Comm* c, size_t bytes) {
    if (bytes <= c->cap_stage) return BSG_OK;
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (c->stage) {
        for (int p = 0; p < c->world && p < kMaxPeers; ++p) {
            for (size_t i = 0; i < c->bufs.size(); ++i)
                if (c->bufs[i].local == c->stage) {
                    for (int q = 0; q < c->world && q < kMaxPeers; ++q)
                        if (q != c->rank && c->bufs[i].peer[q]) cudaIpcCloseMemHandle(c->bufs[i].peer[q]);
                    cudaFree(c->bufs[i].local);
                    c->bufs.erase(c->bufs.begin() + i);
                    break;
                }
            break;
        }
        c->stage = nullptr;
        c->cap_stage = 0;
    }
    SymBuf b;
    if (c->p2p) {
        int mapped = 0;
        int rc = sym_alloc(c, bytes * 2, &b, &mapped);
        if (rc) return rc;
        if (!mapped) return bsg_set_last_error_internal(BSG_ERR_COMM, "a rank could not map a peer's buffer (CUDA IPC)");
    } else {
        b.bytes = (bytes * 2 + kSymAlign - 1) / kSymAlign * kSymAlign;
        CU_TRY(cudaMalloc(reinterpret_cast<void**>(&b.local), b.bytes));
        b.peer[c->rank] = b.local;
    }
    c->bufs.push_back(b);
    c->stage = b.local;
    c->cap_stage = b.bytes;
    return BSG_OK;
}

extern "C" int bsg_or_reduce(bsg_ctx* ctx, uint64_t* words, uint64_t n_words) {
    Comm* c = comm_of(ctx);
    if (!ctx || (n_words && !words)) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_or_reduce: NULL argument");
    if (!c) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_comm_init has not been called");
    if (n_words == 0) return BSG_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CU_TRY(cudaSetDevice(c->device));
    int rc = ensure_stage(c, n_words * 8 + 16);
    if (rc) return rc;
    uint64_t* d = reinterpret_cast<uint64_t*>(c->stage);
    CU_TRY(cudaMemcpyAsync(d, words, n_words * 8, cudaMemcpyHostToDevice, c->stream));
    rc = or_reduce_on(c, d, n_words, c->stream);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(words, d, n_words * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return BSG_OK;
}

// --------------------------------------------------------------- all-gather ---
static int allgather_on(Comm* c, const uint64_t* d_local, uint64_t n_words, uint64_t* d_all, cudaStream_t s) {
    const uint64_t W = static_cast<uint64_t>(c->world);
    c->nvlink_bytes = (W - 1) * n_words * 8;
    if (c->p2p) {
        PeerPtrs all;
        if (!peers_of(c, d_all, n_words * W * 8, &all))
            return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_allgather_masks_device: d_all must be memory from bsg_comm_alloc");
        const Sig sig = make_sig(c);
        const int blocks = static_cast<int>(std::max<uint64_t>(1, std::min<uint64_t>(148, (n_words + 255) / 256)));
        allgather_p2p_kernel<<<blocks, 256, 0, s>>>(d_local, n_words, all, sig);
        CU_TRY(cudaGetLastError());
        return BSG_OK;
    }
    NCCL_TRY(g_nccl.AllGather(d_local, d_all, n_words, ncclUint64, c->comm, s));
    return BSG_OK;
}

extern "C" int bsg_allgather_masks_device(bsg_ctx* ctx, const uint64_t* d_local, uint64_t n_words, uint64_t* d_all) {
    Comm* c = comm_of(ctx);
    if (!c || (n_words && (!d_local || !d_all)))
        return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_allgather_masks_device: bad argument / no communicator");
    if (n_words == 0) return BSG_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CU_TRY(cudaSetDevice(c->device));
    return allgather_on(c, d_local, n_words, d_all, static_cast<cudaStream_t>(bsg_ctx_stream_internal(ctx)));
}

extern "C" int bsg_allgather_masks(bsg_ctx* ctx, const uint64_t* local, uint64_t n_words, uint64_t* all) {
    Comm* c = comm_of(ctx);
    if (!ctx || (n_words && (!local || !all))) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_allgather_masks: NULL argument");
    if (!c) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_comm_init has not been called");
    if (n_words == 0) return BSG_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CU_TRY(cudaSetDevice(c->device));
    const uint64_t W = static_cast<uint64_t>(c->world);
    int rc = ensure_stage(c, (W + 1) * n_words * 8 + 16);
    if (rc) return rc;
    uint64_t* d_all = reinterpret_cast<uint64_t*>(c->stage);
    uint64_t* d_local = d_all + W * n_words;
    CU_TRY(cudaMemcpyAsync(d_local, local, n_words * 8, cudaMemcpyHostToDevice, c->stream));
    rc = allgather_on(c, d_local, n_words, d_all, c->stream);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(all, d_all, n_words * W * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return BSG_OK;
}
