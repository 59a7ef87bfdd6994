// Multi-GPU exchanges (one process per GPU), NCCL over NVLink 5 / NVSwitch.
//
// The probe and the block-level build shard by file with no exchange; the two
// real exchange steps of the path are:
//   bsg_or_reduce       bitwise OR of equal-(m,k) partial file-level bitsets built
//                       from disjoint shards of one file's entries.  NCCL has no
//                       OR reduction, so it is an all-to-all of 1/W slices
//                       (grouped ncclSend/ncclRecv) + a local OR kernel + an
//                       all-gather of the reduced slices (bandwidth-optimal).
//   bsg_allgather_masks all-gather of the per-rank candidate masks.
// libnccl is dlopen'ed so single-GPU users carry no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
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
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
}

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    int device = 0;
    cudaStream_t stream = nullptr;
};

thread_local std::string t_comm_err;

}  // namespace

// api.cu owns bsg_ctx; these accessors keep this file free of its layout.
extern "C" int bsg_ctx_device_internal(bsg_ctx* ctx);
extern "C" void** bsg_ctx_comm_slot_internal(bsg_ctx* ctx);
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
    c->device = bsg_ctx_device_internal(ctx);
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id, 128);
    NCCL_TRY(g_nccl.CommInitRank(&c->comm, world, id, rank));
    *slot = c;
    return BSG_OK;
}

extern "C" void bsg_comm_destroy_internal(void* comm) {
    Comm* c = static_cast<Comm*>(comm);
    if (!c) return;
    if (c->comm && g_nccl.handle) g_nccl.CommDestroy(c->comm);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int bsg_or_reduce(bsg_ctx* ctx, uint64_t* words, uint64_t n_words) {
    if (!ctx || (n_words && !words)) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_or_reduce: NULL argument");
    Comm* c = static_cast<Comm*>(*bsg_ctx_comm_slot_internal(ctx));
    if (!c) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_comm_init has not been called");
    if (n_words == 0) return BSG_OK;
    CU_TRY(cudaSetDevice(c->device));
    const uint64_t W = static_cast<uint64_t>(c->world);
    const uint64_t sl = (n_words + W - 1) / W;  // slice words
    uint64_t *d_full = nullptr, *d_recv = nullptr;
    CU_TRY(cudaMalloc(reinterpret_cast<void**>(&d_full), sl * W * 8));
    CU_TRY(cudaMalloc(reinterpret_cast<void**>(&d_recv), sl * W * 8));
    int rc = BSG_OK;
    auto body = [&]() -> int {
        CU_TRY(cudaMemsetAsync(d_full, 0, sl * W * 8, c->stream));
        CU_TRY(cudaMemcpyAsync(d_full, words, n_words * 8, cudaMemcpyHostToDevice, c->stream));
        if (W > 1) {
            // all-to-all: slice j of every rank lands on rank j
            NCCL_TRY(g_nccl.GroupStart());
            for (uint64_t p = 0; p < W; ++p) {
                NCCL_TRY(g_nccl.Send(d_full + p * sl, sl, ncclUint64, static_cast<int>(p), c->comm, c->stream));
                NCCL_TRY(g_nccl.Recv(d_recv + p * sl, sl, ncclUint64, static_cast<int>(p), c->comm, c->stream));
            }
            NCCL_TRY(g_nccl.GroupEnd());
            // OR the W received copies of my slice
            uint64_t* mine = d_recv + static_cast<uint64_t>(c->rank) * sl;
            for (uint64_t p = 0; p < W; ++p)
                if (p != static_cast<uint64_t>(c->rank)) CU_TRY(bsg::launch_or_words(mine, d_recv + p * sl, sl, c->stream));
            NCCL_TRY(g_nccl.AllGather(mine, d_full, sl, ncclUint64, c->comm, c->stream));
        }
        CU_TRY(cudaMemcpyAsync(words, d_full, n_words * 8, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        return BSG_OK;
    };
    rc = body();
    cudaFree(d_full);
    cudaFree(d_recv);
    return rc;
}

extern "C" int bsg_allgather_masks(bsg_ctx* ctx, const uint64_t* local, uint64_t n_words, uint64_t* all) {
    if (!ctx || (n_words && (!local || !all))) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_allgather_masks: NULL argument");
    Comm* c = static_cast<Comm*>(*bsg_ctx_comm_slot_internal(ctx));
    if (!c) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_comm_init has not been called");
    if (n_words == 0) return BSG_OK;
    CU_TRY(cudaSetDevice(c->device));
    const uint64_t W = static_cast<uint64_t>(c->world);
    uint64_t *d_local = nullptr, *d_all = nullptr;
    CU_TRY(cudaMalloc(reinterpret_cast<void**>(&d_local), n_words * 8));
    CU_TRY(cudaMalloc(reinterpret_cast<void**>(&d_all), n_words * W * 8));
    auto body = [&]() -> int {
        CU_TRY(cudaMemcpyAsync(d_local, local, n_words * 8, cudaMemcpyHostToDevice, c->stream));
        NCCL_TRY(g_nccl.AllGather(d_local, d_all, n_words, ncclUint64, c->comm, c->stream));
        CU_TRY(cudaMemcpyAsync(all, d_all, n_words * W * 8, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        return BSG_OK;
    };
    int rc = body();
    cudaFree(d_local);
    cudaFree(d_all);
    return rc;
}
