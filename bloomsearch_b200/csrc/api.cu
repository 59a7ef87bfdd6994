// C ABI of libbloomgpu.so (include/bloomgpu.h): context, corpus residency,
// query objects, build and probe entry points.  Host-side logic only — all the
// arithmetic lives in the kernels_*.cu files.  There is deliberately no CPU
// path: every entry point needs a CUDA device.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bsg_internal.h"

using namespace bsg;

// ------------------------------------------------------------------ errors ---
static thread_local std::string t_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(_e == cudaErrorMemoryAllocation ? BSG_ERR_NOMEM : BSG_ERR_CUDA, "%s: %s", #expr, \
                        cudaGetErrorString(_e));                                                    \
    } while (0)

extern "C" int bsg_abi_version(void) { return BSG_ABI_VERSION; }

extern "C" const char* bsg_strerror(int code) {
    switch (code) {
    case BSG_OK: return "ok";
    case BSG_ERR_INVALID: return "invalid argument";
    case BSG_ERR_CUDA: return "CUDA error (no device, or runtime failure)";
    case BSG_ERR_NOMEM: return "out of memory";
    case BSG_ERR_FORMAT: return "malformed filter section";
    case BSG_ERR_UNSUPPORTED: return "unsupported";
    case BSG_ERR_COMM: return "NCCL error";
    default: return "unknown error";
    }
}

extern "C" const char* bsg_last_error(void) { return t_last_error.c_str(); }

// ----------------------------------------------------------------- context ---
struct bsg_ctx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    int cc_major = 0, cc_minor = 0;
    size_t total_mem = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t cur_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::mutex mu;
    std::vector<cudaStream_t> stream_pool;
    std::vector<bsg_query*> scratch_pool;  // reusable per-call query objects for bsg_probe
    void* comm = nullptr;  // bsg_comm.cpp
    uint64_t* d_gather = nullptr;  // bsg_probe_hierarchical_gather: symmetric landing zone + padded local mask
    size_t gather_cap = 0;
    uint64_t* h_gather = nullptr;  // pinned copy-out buffer
    // large host<->device transfers of pageable caller memory (bsg_build): worker threads copy
    // slices through pinned double buffers on their own streams (see staged_copy)
    std::mutex stage_mu;
    std::vector<uint8_t*> stage_pin;       // kStageWorkers * 2 buffers of kStageChunk bytes
    std::vector<cudaStream_t> stage_streams;
    std::vector<cudaEvent_t> stage_events;
    std::vector<cudaStream_t> aux_streams;  // bsg_debug_run_cycle
    std::vector<cudaEvent_t> aux_events;
    cudaEvent_t fork_event = nullptr;
    float last_build_kernel_ms = 0.f;  // profiling: device time of the last bsg_build's kernel
    std::mutex distinct_mu;         // bsg_count_distinct: hash-set tables, grow-only
    void* d_distinct = nullptr;
    size_t distinct_cap = 0;
    uint64_t* d_trace = nullptr;  // profiling timeline (bsg_debug_trace_*), [n_ctas][slots]
    uint32_t trace_slots = 0;
    int probe_warps = 0;   // BSG_PROBE_WARPS override (tuning)
    int max_stages = 0;    // BSG_PROBE_STAGES override (tuning)
    int probe_variant = 7; // BSG_PROBE_VARIANT: 7 = per corpus (default, see staged_variant_for); 6 = probe_tiles; 0 = probe_staged
                           // (one phase); 1..5 = shapes of probe_staged2
    int tiles_shape = 1;   // BSG_TILES_SHAPE: compiled shape of probe_tiles_kernel (kernels_probe_tiles.cu)
    int tile_bytes = 80000;      // BSG_TILE_BYTES: UNIT mode, units are grouped into tiles of about this many bytes
    int tile_units = 8;    // BSG_TILE_UNITS: UNIT mode, at most this many units per tile (<= kTileMaxUnits)
    int tile_mode = 0;     // BSG_TILE_MODE: 0 = choose per corpus, 1 = force UNIT mode, 2 = force KIND mode
    int tile_min_stages = 0;  // BSG_TILE_MIN_STAGES: UNIT mode keeps tiles small enough for a ring of this many stages (0 = 2 for grouped units, 3 for a single unit)
    // BSG_PROBE_TIMING=1: host-side phase times of bsg_probe() (ns sums), printed by bsg_destroy
    int timing = 0;
    std::atomic<uint64_t> t_calls{0}, t_prepare{0}, t_run{0}, t_wait{0}, t_copyout{0};
    int fuse_hash = 1;     // BSG_PROBE_FUSE_HASH: bsg_probe() hashes inside the staged probe kernel when it can
    int spin_wait = 1;     // BSG_PROBE_SPIN: bsg_probe() polls the stream instead of blocking in cudaStreamSynchronize
    std::mutex host_mu;
    std::vector<std::pair<uint8_t*, size_t>> host_bufs;  // bsg_host_alloc: pinned + mapped caller buffers
    int short_circuit = 1; // BSG_PROBE_SHORT_CIRCUIT: mask-only small queries on the gather path stop testing a unit's keys once its expression is decided
    int zero_copy = 1;     // BSG_PROBE_ZEROCOPY: bsg_probe() matrix rows written straight to pinned host memory
    int pdl = 1;           // BSG_PROBE_PDL: programmatic dependent launch of the two-phase probe kernel
    int relax_sleep_ns = 0;  // BSG_PROBE_SLEEP: ns slept between polls of a phase-B warp (measured: no effect)
    int stagger_pct = -1;  // BSG_PROBE_STAGGER: % of the one-stage-per-SM stream time between prologue fills
};

extern "C" void bsg_comm_destroy_internal(void* comm);
extern "C" void bsg_query_free(bsg_query* q);

static cudaStream_t pool_get(bsg_ctx* ctx) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->stream_pool.empty()) {
        cudaStream_t s = ctx->stream_pool.back();
        ctx->stream_pool.pop_back();
        return s;
    }
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    return s;
}
static void pool_put(bsg_ctx* ctx, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->stream_pool.push_back(s);
}

extern "C" int bsg_create(int device, bsg_ctx** out) {
    if (!out) return fail(BSG_ERR_INVALID, "bsg_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(BSG_ERR_CUDA, "no CUDA device available (%s); libbloomgpu has no CPU fallback",
                    e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(BSG_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    CUDA_TRY(cudaSetDevice(device));
    bsg_ctx* ctx = new (std::nothrow) bsg_ctx();
    if (!ctx) return fail(BSG_ERR_NOMEM, "ctx alloc");
    ctx->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->total_mem = prop.totalGlobalMem;
    if (prop.major < 10) {
        delete ctx;
        return fail(BSG_ERR_UNSUPPORTED, "device compute capability %d.%d < 10.0: this library is sm_100a only",
                    prop.major, prop.minor);
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->cur_stream = ctx->own_stream;
    CUDA_TRY(cudaEventCreate(&ctx->ev0));
    CUDA_TRY(cudaEventCreate(&ctx->ev1));
    CUDA_TRY(probe_staged_configure(ctx->max_smem_optin));
    CUDA_TRY(probe_tiles_configure(ctx->max_smem_optin));
    CUDA_TRY(build_configure(ctx->max_smem_optin));
    CUDA_TRY(sections_configure());
    // BSG_L2_FETCH = 32 | 64 | 128: L2 fetch granularity hint (device-wide).  The gather kernel needs 4 bytes of every
    // 32-byte sector it touches; a coarser fetch only multiplies its DRAM traffic.
    if (const char* w = getenv("BSG_L2_FETCH")) {
        const int g = atoi(w);
        if (g == 32 || g == 64 || g == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, static_cast<size_t>(g));
    }
    if (const char* w = getenv("BSG_PROBE_WARPS")) ctx->probe_warps = atoi(w);
    if (const char* w = getenv("BSG_PROBE_STAGES")) ctx->max_stages = atoi(w);
    if (const char* w = getenv("BSG_PROBE_STAGGER")) ctx->stagger_pct = atoi(w);
    if (const char* w = getenv("BSG_PROBE_VARIANT")) ctx->probe_variant = std::min(7, std::max(0, atoi(w)));
    if (const char* w = getenv("BSG_TILES_SHAPE")) ctx->tiles_shape = std::min(probe_tiles_n_shapes() - 1, std::max(0, atoi(w)));
    if (const char* w = getenv("BSG_TILE_BYTES")) ctx->tile_bytes = std::max(1024, atoi(w));
    if (const char* w = getenv("BSG_TILE_UNITS")) ctx->tile_units = std::min<int>(kTileMaxUnits, std::max(1, atoi(w)));
    if (const char* w = getenv("BSG_TILE_MODE")) ctx->tile_mode = std::min(2, std::max(0, atoi(w)));
    if (const char* w = getenv("BSG_TILE_MIN_STAGES")) ctx->tile_min_stages = std::min(8, std::max(0, atoi(w)));
    if (const char* w = getenv("BSG_PROBE_PDL")) ctx->pdl = atoi(w) != 0;
    if (const char* w = getenv("BSG_PROBE_TIMING")) ctx->timing = atoi(w);
    if (const char* w = getenv("BSG_PROBE_FUSE_HASH")) ctx->fuse_hash = atoi(w) != 0;
    if (const char* w = getenv("BSG_PROBE_ZEROCOPY")) ctx->zero_copy = atoi(w) != 0;
    if (const char* w = getenv("BSG_PROBE_SHORT_CIRCUIT")) ctx->short_circuit = atoi(w) != 0;
    if (const char* w = getenv("BSG_PROBE_SPIN")) ctx->spin_wait = atoi(w) != 0;
    if (const char* w = getenv("BSG_PROBE_SLEEP")) ctx->relax_sleep_ns = std::max(0, atoi(w));
    *out = ctx;
    return BSG_OK;
}

extern "C" void bsg_destroy(bsg_ctx* ctx) {
    if (!ctx) return;
    if (ctx->timing && ctx->t_calls) {
        const double n = static_cast<double>(ctx->t_calls.load()) * 1e3;
        fprintf(stderr, "[bsg] bsg_probe host phases over %llu calls (us/call): prepare+enqueue H2D/hash %.1f, enqueue probe %.1f, "
                        "enqueue D2H + wait %.1f, copy out %.1f\n", (unsigned long long)ctx->t_calls.load(),
                ctx->t_prepare / n, ctx->t_run / n, ctx->t_wait / n, ctx->t_copyout / n);
    }
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
    if (ctx->comm) bsg_comm_destroy_internal(ctx->comm);  // also frees the symmetric buffers (d_gather)
    for (cudaStream_t s : ctx->stream_pool) cudaStreamDestroy(s);
    for (bsg_query* q : ctx->scratch_pool) bsg_query_free(q);
    for (auto& hb : ctx->host_bufs) cudaFreeHost(hb.first);
    for (uint8_t* p : ctx->stage_pin) cudaFreeHost(p);
    for (cudaStream_t s : ctx->stage_streams) cudaStreamDestroy(s);
    for (cudaEvent_t e : ctx->stage_events) cudaEventDestroy(e);
    for (cudaStream_t s : ctx->aux_streams) cudaStreamDestroy(s);
    for (cudaEvent_t e : ctx->aux_events) cudaEventDestroy(e);
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    cudaFree(ctx->d_trace);
    cudaFree(ctx->d_distinct);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" int bsg_set_stream(bsg_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(BSG_ERR_INVALID, "ctx is NULL");
    ctx->cur_stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return BSG_OK;
}

extern "C" int bsg_synchronize(bsg_ctx* ctx) {
    if (!ctx) return fail(BSG_ERR_INVALID, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->cur_stream));
    return BSG_OK;
}

extern "C" int bsg_device_info(bsg_ctx* ctx, int* sm_count, size_t* smem_optin, size_t* total_mem, int* cc_major,
                               int* cc_minor) {
    if (!ctx) return fail(BSG_ERR_INVALID, "ctx is NULL");
    if (sm_count) *sm_count = ctx->sm_count;
    if (smem_optin) *smem_optin = static_cast<size_t>(ctx->max_smem_optin);
    if (total_mem) *total_mem = ctx->total_mem;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    return BSG_OK;
}

// Pinned, device-mapped host memory for result buffers: bsg_probe() lets its kernels write a matrix that lives
// in such a buffer directly (no staging copy at the end of the call).
extern "C" int bsg_host_alloc(bsg_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return fail(BSG_ERR_INVALID, "NULL argument");
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    void* p = nullptr;
    CUDA_TRY(cudaHostAlloc(&p, std::max<size_t>(bytes, 16), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(p, 0, std::max<size_t>(bytes, 16));
    std::lock_guard<std::mutex> lk(ctx->host_mu);
    ctx->host_bufs.emplace_back(static_cast<uint8_t*>(p), std::max<size_t>(bytes, 16));
    *out = p;
    return BSG_OK;
}

extern "C" int bsg_host_free(bsg_ctx* ctx, void* ptr) {
    if (!ctx) return fail(BSG_ERR_INVALID, "NULL argument");
    if (!ptr) return BSG_OK;
    std::lock_guard<std::mutex> lk(ctx->host_mu);
    for (size_t i = 0; i < ctx->host_bufs.size(); ++i)
        if (ctx->host_bufs[i].first == ptr) {
            cudaFreeHost(ptr);
            ctx->host_bufs.erase(ctx->host_bufs.begin() + i);
            return BSG_OK;
        }
    return fail(BSG_ERR_INVALID, "bsg_host_free: not a bsg_host_alloc buffer");
}

static uint32_t* host_buf_device_ptr(bsg_ctx* ctx, const void* p, size_t bytes) {
    const uint8_t* p8 = static_cast<const uint8_t*>(p);
    {
        std::lock_guard<std::mutex> lk(ctx->host_mu);
        bool found = false;
        for (auto& hb : ctx->host_bufs)
            if (p8 >= hb.first && p8 + bytes <= hb.first + hb.second) { found = true; break; }
        if (!found) return nullptr;
    }
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, const_cast<void*>(p), 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return static_cast<uint32_t*>(d);
}

// wait for a stream without parking the thread in the driver (a 15-20 us kernel is over before a blocked
// thread is rescheduled); falls back to the blocking call after ~2 ms
static cudaError_t stream_wait(bsg_ctx* ctx, cudaStream_t s) {
    if (ctx && ctx->spin_wait) {
        const auto t0 = std::chrono::steady_clock::now();
        for (uint32_t i = 0;; ++i) {
            const cudaError_t e = cudaStreamQuery(s);
            if (e != cudaErrorNotReady) return e;
            if ((i & 63u) == 63u && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) break;
        }
    }
    return cudaStreamSynchronize(s);
}

extern "C" int bsg_timer_begin(bsg_ctx* ctx) {
    if (!ctx) return fail(BSG_ERR_INVALID, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->cur_stream));
    return BSG_OK;
}

extern "C" int bsg_timer_end(bsg_ctx* ctx, float* elapsed_ms) {
    if (!ctx || !elapsed_ms) return fail(BSG_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->cur_stream));
    CUDA_TRY(cudaEventSynchronize(ctx->ev1));
    CUDA_TRY(cudaEventElapsedTime(elapsed_ms, ctx->ev0, ctx->ev1));
    return BSG_OK;
}

// ------------------------------------------------------------------ sizing ---
extern "C" void bsg_estimate(uint64_t n, double fpr, uint64_t* m, uint64_t* k) {
    // bloom.EstimateParameters (float64) followed by New()'s clamp to >= 1.  A Go host
    // calls the library itself and passes integers; this helper is for other hosts.
    const double ln2 = std::log(2.0);
    const double md = std::ceil(-1.0 * static_cast<double>(n) * std::log(fpr) / std::pow(ln2, 2.0));
    const uint64_t mm = static_cast<uint64_t>(md);
    const double kd = std::ceil(ln2 * static_cast<double>(mm) / static_cast<double>(n));
    const uint64_t kk = static_cast<uint64_t>(kd);
    if (m) *m = mm < 1 ? 1 : mm;
    if (k) *k = kk < 1 ? 1 : kk;
}

// ----------------------------------------------------------- small helpers ---
namespace {

template <typename T>
struct DevBuf {  // RAII device allocation
    T* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(reinterpret_cast<void**>(&p), std::max<size_t>(n, 1) * sizeof(T)); }
    T* release() { T* r = p; p = nullptr; return r; }
};

constexpr uint64_t kMaxM = 1ull << 62;
constexpr uint64_t kKeyPad = 16;  // WordReader may read up to the 8-byte boundary + one word past a key

inline uint64_t words_for(uint64_t m) { return (m + 63) >> 6; }
inline uint64_t even_up(uint64_t w) { return (w + 1) & ~1ull; }

// key_off must be monotone with keys < 4 GiB; checked with a few host threads for large batches
// (a 40 M-key build would otherwise spend ~40 ms in this loop).  Returns the first bad index or
// n_keys when everything is fine.
uint64_t first_bad_offset(const uint64_t* key_off, uint64_t n_keys) {
    auto scan = [&](uint64_t lo, uint64_t hi) -> uint64_t {
        for (uint64_t i = lo; i < hi; ++i)
            if (key_off[i + 1] < key_off[i] || key_off[i + 1] - key_off[i] > 0xffffffffull) return i;
        return n_keys;
    };
    if (n_keys < (1u << 20)) return scan(0, n_keys);
    const int T = 8;
    std::vector<uint64_t> bad(T, n_keys);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t] { bad[t] = scan(n_keys * t / T, n_keys * (t + 1) / T); });
    for (auto& x : th) x.join();
    uint64_t r = n_keys;
    for (uint64_t b : bad) r = std::min(r, b);
    return r;
}

int validate_filter(const bsg_filter_desc& d, uint64_t n_words, const char* what, uint64_t idx) {
    if (d.m == 0) return BSG_OK;
    if (d.m > kMaxM) return fail(BSG_ERR_INVALID, "%s %llu: m=%llu exceeds 2^62", what, (unsigned long long)idx,
                                 (unsigned long long)d.m);
    if (d.k == 0 || d.k > 0x7fffffffull)
        return fail(BSG_ERR_INVALID, "%s %llu: k=%llu out of range", what, (unsigned long long)idx,
                    (unsigned long long)d.k);
    const uint64_t nw = words_for(d.m);
    if (nw > 0xffffffffull) return fail(BSG_ERR_INVALID, "%s %llu: filter too large", what, (unsigned long long)idx);
    if (d.word_off > n_words || nw > n_words - d.word_off)
        return fail(BSG_ERR_INVALID, "%s %llu: words [%llu,+%llu) outside the words array (%llu)", what,
                    (unsigned long long)idx, (unsigned long long)d.word_off, (unsigned long long)nw,
                    (unsigned long long)n_words);
    return BSG_OK;
}

}  // namespace

// ------------------------------------------------------------ staged copies ---
// cudaMemcpy from/to pageable memory runs at ~5 GB/s; the caller's buffers (Go heap) cannot be
// pinned in place.  staged_copy moves a large array with kStageWorkers host threads, each copying
// its chunks through two pinned buffers (memcpy overlapped with the previous chunk's DMA) on its
// own stream.  Blocks until the whole transfer is complete.
namespace {
constexpr int kStageWorkers = 12;
constexpr size_t kStageChunk = 4u << 20;
constexpr size_t kStageThreshold = 4u << 20;  // below this a plain cudaMemcpyAsync is used

cudaError_t stage_init(bsg_ctx* ctx) {
    if (!ctx->stage_pin.empty()) return cudaSuccess;
    for (int i = 0; i < kStageWorkers * 2; ++i) {
        uint8_t* p = nullptr;
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&p), kStageChunk, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        ctx->stage_pin.push_back(p);
        cudaEvent_t ev = nullptr;
        e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        ctx->stage_events.push_back(ev);
    }
    for (int i = 0; i < kStageWorkers; ++i) {
        cudaStream_t st = nullptr;
        cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        if (e != cudaSuccess) return e;
        ctx->stage_streams.push_back(st);
    }
    return cudaSuccess;
}

// to_device: dev <- host ; else host <- dev.  Caller holds ctx->stage_mu.
cudaError_t staged_copy(bsg_ctx* ctx, void* dev, void* host, size_t bytes, bool to_device) {
    if (bytes == 0) return cudaSuccess;
    cudaError_t e = stage_init(ctx);
    if (e != cudaSuccess) return e;
    const size_t n_chunks = (bytes + kStageChunk - 1) / kStageChunk;
    std::vector<cudaError_t> errs(kStageWorkers, cudaSuccess);
    auto worker = [&](int w) {
        cudaSetDevice(ctx->device);
        cudaStream_t st = ctx->stage_streams[w];
        int j = 0;
        size_t pending_off[2] = {0, 0}, pending_len[2] = {0, 0};
        bool pending[2] = {false, false};
        for (size_t c = w; c < n_chunks; c += kStageWorkers, ++j) {
            const int b = j & 1;
            uint8_t* pin = ctx->stage_pin[w * 2 + b];
            cudaEvent_t ev = ctx->stage_events[w * 2 + b];
            const size_t off = c * kStageChunk, len = std::min(kStageChunk, bytes - off);
            if (pending[b]) {  // this buffer's previous DMA must be done before it is reused
                cudaError_t r = cudaEventSynchronize(ev);
                if (r != cudaSuccess) { errs[w] = r; return; }
                if (!to_device) memcpy(static_cast<uint8_t*>(host) + pending_off[b], pin, pending_len[b]);
                pending[b] = false;
            }
            cudaError_t r;
            if (to_device) {
                memcpy(pin, static_cast<const uint8_t*>(host) + off, len);
                r = cudaMemcpyAsync(static_cast<uint8_t*>(dev) + off, pin, len, cudaMemcpyHostToDevice, st);
            } else {
                r = cudaMemcpyAsync(pin, static_cast<const uint8_t*>(dev) + off, len, cudaMemcpyDeviceToHost, st);
            }
            if (r == cudaSuccess) r = cudaEventRecord(ev, st);
            if (r != cudaSuccess) { errs[w] = r; return; }
            pending[b] = true;
            pending_off[b] = off;
            pending_len[b] = len;
        }
        for (int b = 0; b < 2; ++b) {
            // drain in issue order: the older buffer first
            const int bb = (j + b) & 1;
            if (!pending[bb]) continue;
            cudaError_t r = cudaEventSynchronize(ctx->stage_events[w * 2 + bb]);
            if (r != cudaSuccess) { errs[w] = r; return; }
            if (!to_device) memcpy(static_cast<uint8_t*>(host) + pending_off[bb], ctx->stage_pin[w * 2 + bb], pending_len[bb]);
        }
    };
    std::vector<std::thread> th;
    const int n_workers = static_cast<int>(std::min<size_t>(kStageWorkers, n_chunks));
    for (int w = 1; w < n_workers; ++w) th.emplace_back(worker, w);
    worker(0);
    for (auto& t : th) t.join();
    for (cudaError_t r : errs)
        if (r != cudaSuccess) return r;
    return cudaSuccess;
}

bool is_pinned_caller_buffer(bsg_ctx* ctx, const void* p, size_t bytes) {
    const uint8_t* p8 = static_cast<const uint8_t*>(p);
    std::lock_guard<std::mutex> lk(ctx->host_mu);
    for (auto& hb : ctx->host_bufs)
        if (p8 >= hb.first && p8 + bytes <= hb.first + hb.second) return true;
    return false;
}

// host -> device on stream s for small arrays and for pinned caller buffers (bsg_host_alloc: one DMA at PCIe
// speed), staged through pinned double buffers (and synchronous) for large pageable ones
cudaError_t upload(bsg_ctx* ctx, void* dev, const void* host, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return cudaSuccess;
    if (bytes < kStageThreshold || is_pinned_caller_buffer(ctx, host, bytes))
        return cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s);
    return staged_copy(ctx, dev, const_cast<void*>(host), bytes, true);
}
}  // namespace

// -------------------------------------------------------------------- hash ---
extern "C" int bsg_hash_keys(bsg_ctx* ctx, const uint8_t* keys, const uint64_t* key_off, uint64_t n_keys,
                             uint64_t* out_hashes) {
    if (!ctx || !key_off || !out_hashes) return fail(BSG_ERR_INVALID, "NULL argument");
    if (n_keys == 0) return BSG_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t nbytes = key_off[n_keys];
    if (nbytes && !keys) return fail(BSG_ERR_INVALID, "keys is NULL");
    {
        const uint64_t bad = first_bad_offset(key_off, n_keys);
        if (bad != n_keys) return fail(BSG_ERR_INVALID, "key_off not monotone at %llu", (unsigned long long)bad);
    }
    cudaStream_t s = pool_get(ctx);
    if (!s) return fail(BSG_ERR_CUDA, "stream create failed");
    DevBuf<uint8_t> d_keys;
    DevBuf<uint64_t> d_off, d_h;
    int rc = BSG_OK;
    do {
        if (d_keys.alloc(nbytes + kKeyPad) != cudaSuccess || d_off.alloc(n_keys + 1) != cudaSuccess ||
            d_h.alloc(n_keys * 4) != cudaSuccess) { rc = fail(BSG_ERR_NOMEM, "device alloc"); break; }
        cudaError_t e = cudaSuccess;
        std::unique_lock<std::mutex> stage_lk(ctx->stage_mu);  // one staged transfer set at a time
        // memsets first: the staged uploads below block, the small async copies after them do not
        e = cudaMemsetAsync(d_keys.p + nbytes, 0, kKeyPad, s);
        if (e == cudaSuccess) e = upload(ctx, d_keys.p, keys, nbytes, s);
        if (e == cudaSuccess) e = upload(ctx, d_off.p, key_off, (n_keys + 1) * 8, s);
        if (e == cudaSuccess) e = launch_hash_keys(d_keys.p, d_off.p, n_keys, d_h.p, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_hashes, d_h.p, n_keys * 32, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = fail(BSG_ERR_CUDA, "bsg_hash_keys: %s", cudaGetErrorString(e));
    } while (0);
    pool_put(ctx, s);
    return rc;
}

// ------------------------------------------------------------------- build ---
namespace {
// Validated launch plan of a build: filter descriptors over the caller's out_words layout, and the key
// groups split so that one CTA never owns more than kSplit keys (sub-groups share the filters; the staged
// bitset is merged with RED.OR).  group_begin is a CSR array, so sub-group boundaries stay contiguous.
struct BuildPlan {
    std::vector<BuildFilter> bf;
    std::vector<uint64_t> gbe;
    std::vector<uint32_t> gf, gf2;
    uint32_t n_sub = 0;
    uint32_t smem_cap = 0;
};
int plan_build(bsg_ctx* ctx, uint64_t n_keys, const uint64_t* group_begin, uint32_t n_groups, const uint32_t* group_filter,
               const uint32_t* group_filter2, const bsg_filter_desc* desc, uint32_t n_filters, uint64_t n_words,
               BuildPlan& P) {
    P.bf.resize(n_filters);
    for (uint32_t f = 0; f < n_filters; ++f) {
        if (desc[f].m == 0) return fail(BSG_ERR_INVALID, "filter %u: m == 0", f);
        int rc = validate_filter(desc[f], n_words, "filter", f);
        if (rc) return rc;
        P.bf[f] = BuildFilter{desc[f].word_off, desc[f].m, reciprocal(desc[f].m), static_cast<uint32_t>(desc[f].k),
                              static_cast<uint32_t>(words_for(desc[f].m))};
    }
    constexpr uint64_t kSplit = 16384;
    P.gbe.reserve(static_cast<size_t>(n_groups) + 1);
    P.gf.reserve(n_groups);
    if (group_filter2) P.gf2.reserve(n_groups);
    for (uint32_t g = 0; g < n_groups; ++g) {
        const uint64_t b = group_begin[g], e = group_begin[g + 1];
        if (e < b || e > n_keys) return fail(BSG_ERR_INVALID, "group %u: key range [%llu,%llu) invalid", g,
                                             (unsigned long long)b, (unsigned long long)e);
        if (group_filter[g] >= n_filters) return fail(BSG_ERR_INVALID, "group %u: filter id out of range", g);
        if (group_filter2 && group_filter2[g] != BSG_NO_FILTER && group_filter2[g] >= n_filters)
            return fail(BSG_ERR_INVALID, "group %u: secondary filter id out of range", g);
        for (uint64_t s0 = b; s0 < e; s0 += kSplit) {
            P.gbe.push_back(s0);
            P.gf.push_back(group_filter[g]);
            if (group_filter2) P.gf2.push_back(group_filter2[g]);
        }
    }
    P.gbe.push_back(n_groups ? group_begin[n_groups] : 0);
    if (P.gf.size() > 0x7fffffffull) return fail(BSG_ERR_INVALID, "too many groups");
    P.n_sub = static_cast<uint32_t>(P.gf.size());
    // shared-memory staging capacity: largest primary filter that still lets 2 CTAs share an SM
    const uint32_t cap_limit = static_cast<uint32_t>(std::min<int>(ctx->max_smem_optin, 100 * 1024));
    P.smem_cap = 0;
    for (uint32_t g = 0; g < P.n_sub; ++g) {
        const uint64_t bytes = static_cast<uint64_t>(P.bf[P.gf[g]].nwords) * 8;
        if (bytes <= cap_limit) P.smem_cap = std::max<uint32_t>(P.smem_cap, static_cast<uint32_t>((bytes + 15) & ~15ull));
    }
    return BSG_OK;
}
}  // namespace

extern "C" int bsg_build(bsg_ctx* ctx, const uint8_t* keys, const uint64_t* key_off, uint64_t n_keys,
                         const uint64_t* group_begin, uint32_t n_groups, const uint32_t* group_filter,
                         const uint32_t* group_filter2, const bsg_filter_desc* desc, uint32_t n_filters,
                         uint64_t* out_words, uint64_t n_words) {
    if (!ctx || !key_off || !desc || !out_words || (n_groups && (!group_begin || !group_filter)))
        return fail(BSG_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t nbytes = n_keys ? key_off[n_keys] : 0;
    if (nbytes && !keys) return fail(BSG_ERR_INVALID, "keys is NULL");
    {
        const uint64_t bad = first_bad_offset(key_off, n_keys);
        if (bad != n_keys) return fail(BSG_ERR_INVALID, "key_off not monotone at %llu", (unsigned long long)bad);
    }
    BuildPlan P;
    {
        int prc = plan_build(ctx, n_keys, group_begin, n_groups, group_filter, group_filter2, desc, n_filters, n_words, P);
        if (prc) return prc;
    }
    const std::vector<BuildFilter>& bf = P.bf;
    const std::vector<uint64_t>& gbe = P.gbe;
    const std::vector<uint32_t>&gf = P.gf, &gf2 = P.gf2;
    const uint32_t n_sub = P.n_sub, smem_cap = P.smem_cap;

    cudaStream_t s = pool_get(ctx);
    if (!s) return fail(BSG_ERR_CUDA, "stream create failed");
    DevBuf<uint8_t> d_keys;
    DevBuf<uint64_t> d_off, d_gb, d_out;
    DevBuf<uint32_t> d_gf, d_gf2;
    DevBuf<BuildFilter> d_bf;
    int rc = BSG_OK;
    do {
        if (d_keys.alloc(nbytes + kKeyPad) != cudaSuccess || d_off.alloc(n_keys + 1) != cudaSuccess ||
            d_gb.alloc(gbe.size()) != cudaSuccess || d_out.alloc(n_words) != cudaSuccess ||
            d_gf.alloc(n_sub) != cudaSuccess || d_bf.alloc(n_filters) != cudaSuccess ||
            (group_filter2 && d_gf2.alloc(n_sub) != cudaSuccess)) { rc = fail(BSG_ERR_NOMEM, "device alloc"); break; }
        cudaError_t e = cudaSuccess;
        std::unique_lock<std::mutex> stage_lk(ctx->stage_mu);  // one staged transfer set at a time
        // memsets first: the staged uploads below block, the small async copies after them do not
        e = cudaMemsetAsync(d_keys.p + nbytes, 0, kKeyPad, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_out.p, 0, std::max<uint64_t>(n_words, 1) * 8, s);
        if (e == cudaSuccess) e = upload(ctx, d_keys.p, keys, nbytes, s);
        if (e == cudaSuccess) e = upload(ctx, d_off.p, key_off, (n_keys + 1) * 8, s);
        if (e == cudaSuccess && n_sub) e = cudaMemcpyAsync(d_gb.p, gbe.data(), gbe.size() * 8, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && n_sub) e = cudaMemcpyAsync(d_gf.p, gf.data(), n_sub * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && n_sub && group_filter2)
            e = cudaMemcpyAsync(d_gf2.p, gf2.data(), n_sub * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && n_filters)
            e = cudaMemcpyAsync(d_bf.p, bf.data(), n_filters * sizeof(BuildFilter), cudaMemcpyHostToDevice, s);
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (e == cudaSuccess) e = cudaEventCreate(&ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&ev1);
        if (e == cudaSuccess) e = cudaEventRecord(ev0, s);
        if (e == cudaSuccess)
            e = launch_build(d_keys.p, d_off.p, d_gb.p, n_sub, d_gf.p, group_filter2 ? d_gf2.p : nullptr, d_bf.p,
                             d_out.p, smem_cap, s);
        if (e == cudaSuccess) e = cudaEventRecord(ev1, s);
        if (e == cudaSuccess) e = cudaEventSynchronize(ev1);
        if (e == cudaSuccess) cudaEventElapsedTime(&ctx->last_build_kernel_ms, ev0, ev1);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        // ev1 was synchronised above: the kernel is done, d_out is final
        if (e == cudaSuccess && n_words) {
            if (n_words * 8 < kStageThreshold) {
                e = cudaMemcpyAsync(out_words, d_out.p, n_words * 8, cudaMemcpyDeviceToHost, s);
                if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            } else {
                e = staged_copy(ctx, d_out.p, out_words, n_words * 8, false);
            }
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = fail(BSG_ERR_CUDA, "bsg_build: %s", cudaGetErrorString(e));
    } while (0);
    pool_put(ctx, s);
    return rc;
}

// Fused field::token build.  Same contract as bsg_build, but entry i of a group
// is (pair_path[i], pair_token[i]) into one string table and its key is path + "::" + token.
extern "C" int bsg_build_fieldtokens(bsg_ctx* ctx, const uint8_t* strings, const uint64_t* str_off, uint64_t n_strings,
                                     const uint32_t* pair_path, const uint32_t* pair_token, uint64_t n_pairs,
                                     const uint64_t* group_begin, uint32_t n_groups, const uint32_t* group_filter,
                                     const uint32_t* group_filter2, const bsg_filter_desc* desc, uint32_t n_filters,
                                     uint64_t* out_words, uint64_t n_words) {
    if (!ctx || !str_off || !desc || !out_words || (n_pairs && (!pair_path || !pair_token)) ||
        (n_groups && (!group_begin || !group_filter)))
        return fail(BSG_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t nbytes = n_strings ? str_off[n_strings] : 0;
    if (nbytes && !strings) return fail(BSG_ERR_INVALID, "strings is NULL");
    {
        const uint64_t bad = first_bad_offset(str_off, n_strings);
        if (bad != n_strings) return fail(BSG_ERR_INVALID, "str_off not monotone at %llu", (unsigned long long)bad);
    }
    for (uint64_t i = 0; i < n_pairs; ++i)
        if (pair_path[i] >= n_strings || pair_token[i] >= n_strings)
            return fail(BSG_ERR_INVALID, "pair %llu: string index out of range", (unsigned long long)i);
    std::vector<BuildFilter> bf(n_filters);
    for (uint32_t f = 0; f < n_filters; ++f) {
        if (desc[f].m == 0) return fail(BSG_ERR_INVALID, "filter %u: m == 0", f);
        int rc = validate_filter(desc[f], n_words, "filter", f);
        if (rc) return rc;
        bf[f] = BuildFilter{desc[f].word_off, desc[f].m, reciprocal(desc[f].m), static_cast<uint32_t>(desc[f].k),
                            static_cast<uint32_t>(words_for(desc[f].m))};
    }
    constexpr uint64_t kSplit = 16384;
    std::vector<uint64_t> gbe;
    std::vector<uint32_t> gf, gf2;
    for (uint32_t g = 0; g < n_groups; ++g) {
        const uint64_t b = group_begin[g], e = group_begin[g + 1];
        if (e < b || e > n_pairs) return fail(BSG_ERR_INVALID, "group %u: pair range invalid", g);
        if (group_filter[g] >= n_filters) return fail(BSG_ERR_INVALID, "group %u: filter id out of range", g);
        if (group_filter2 && group_filter2[g] != BSG_NO_FILTER && group_filter2[g] >= n_filters)
            return fail(BSG_ERR_INVALID, "group %u: secondary filter id out of range", g);
        for (uint64_t s0 = b; s0 < e; s0 += kSplit) {
            gbe.push_back(s0);
            gf.push_back(group_filter[g]);
            if (group_filter2) gf2.push_back(group_filter2[g]);
        }
    }
    gbe.push_back(n_groups ? group_begin[n_groups] : 0);
    const uint32_t n_sub = static_cast<uint32_t>(gf.size());
    const uint32_t cap_limit = static_cast<uint32_t>(std::min<int>(ctx->max_smem_optin, 100 * 1024));
    uint32_t smem_cap = 0;
    for (uint32_t g = 0; g < n_sub; ++g) {
        const uint64_t bytes = static_cast<uint64_t>(bf[gf[g]].nwords) * 8;
        if (bytes <= cap_limit) smem_cap = std::max<uint32_t>(smem_cap, static_cast<uint32_t>((bytes + 15) & ~15ull));
    }
    cudaStream_t s = pool_get(ctx);
    if (!s) return fail(BSG_ERR_CUDA, "stream create failed");
    DevBuf<uint8_t> d_str;
    DevBuf<uint64_t> d_off, d_gb, d_out;
    DevBuf<uint32_t> d_pp, d_pt, d_gf, d_gf2;
    DevBuf<BuildFilter> d_bf;
    int rc = BSG_OK;
    do {
        if (d_str.alloc(nbytes + kKeyPad) != cudaSuccess || d_off.alloc(n_strings + 1) != cudaSuccess ||
            d_pp.alloc(n_pairs) != cudaSuccess || d_pt.alloc(n_pairs) != cudaSuccess ||
            d_gb.alloc(gbe.size()) != cudaSuccess || d_out.alloc(n_words) != cudaSuccess ||
            d_gf.alloc(n_sub) != cudaSuccess || d_bf.alloc(n_filters) != cudaSuccess ||
            (group_filter2 && d_gf2.alloc(n_sub) != cudaSuccess)) { rc = fail(BSG_ERR_NOMEM, "device alloc"); break; }
        std::unique_lock<std::mutex> stage_lk(ctx->stage_mu);
        cudaError_t e = cudaMemsetAsync(d_str.p + nbytes, 0, kKeyPad, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_out.p, 0, std::max<uint64_t>(n_words, 1) * 8, s);
        if (e == cudaSuccess) e = upload(ctx, d_str.p, strings, nbytes, s);
        if (e == cudaSuccess) e = upload(ctx, d_off.p, str_off, (n_strings + 1) * 8, s);
        if (e == cudaSuccess) e = upload(ctx, d_pp.p, pair_path, n_pairs * 4, s);
        if (e == cudaSuccess) e = upload(ctx, d_pt.p, pair_token, n_pairs * 4, s);
        if (e == cudaSuccess && n_sub) e = cudaMemcpyAsync(d_gb.p, gbe.data(), gbe.size() * 8, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && n_sub) e = cudaMemcpyAsync(d_gf.p, gf.data(), n_sub * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && n_sub && group_filter2)
            e = cudaMemcpyAsync(d_gf2.p, gf2.data(), n_sub * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && n_filters)
            e = cudaMemcpyAsync(d_bf.p, bf.data(), n_filters * sizeof(BuildFilter), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess)
            e = launch_build_ft(d_str.p, d_off.p, d_pp.p, d_pt.p, d_gb.p, n_sub, d_gf.p, group_filter2 ? d_gf2.p : nullptr,
                                d_bf.p, d_out.p, smem_cap, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e == cudaSuccess && n_words) {
            if (n_words * 8 < kStageThreshold) {
                e = cudaMemcpyAsync(out_words, d_out.p, n_words * 8, cudaMemcpyDeviceToHost, s);
                if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            } else {
                e = staged_copy(ctx, d_out.p, out_words, n_words * 8, false);
            }
        }
        if (e != cudaSuccess) rc = fail(BSG_ERR_CUDA, "bsg_build_fieldtokens: %s", cudaGetErrorString(e));
    } while (0);
    pool_put(ctx, s);
    return rc;
}

// Counts on device-resident keys (shared by bsg_count_distinct and bsg_keyset_count_distinct).
static int count_distinct_device(bsg_ctx* ctx, const uint8_t* d_keys, const uint64_t* d_off, uint64_t n_keys,
                                 const uint64_t* d_gb, uint32_t n_groups, const uint32_t* group_parent, uint32_t n_parents,
                                 uint64_t* out_group_counts, uint64_t* out_parent_counts, cudaStream_t s) {
    // the hash-set tables (GBs for a flush-sized batch) are kept by the ctx and only grow: allocating and freeing
    // them per call cost more than the counting kernel; one count at a time per ctx (the flush worker's pace)
    std::lock_guard<std::mutex> distinct_lk(ctx->distinct_mu);
    const size_t need = count_distinct_scratch_bytes(n_keys);
    if (need > ctx->distinct_cap) {
        cudaFree(ctx->d_distinct);
        ctx->d_distinct = nullptr;
        ctx->distinct_cap = 0;
        if (cudaMalloc(&ctx->d_distinct, need) != cudaSuccess) { cudaGetLastError(); return fail(BSG_ERR_NOMEM, "device alloc (%zu bytes of hash-set tables)", need); }
        ctx->distinct_cap = need;
    }
    struct { void* p; } d_em{ctx->d_distinct};
    DevBuf<uint32_t> d_gp;
    DevBuf<unsigned long long> d_gc, d_pc;
    if (d_gc.alloc(n_groups) != cudaSuccess ||
        (group_parent && (d_gp.alloc(n_groups) != cudaSuccess || d_pc.alloc(n_parents) != cudaSuccess)))
        return fail(BSG_ERR_NOMEM, "device alloc");
    cudaError_t e = cudaMemsetAsync(d_gc.p, 0, n_groups * 8, s);
    if (e == cudaSuccess && group_parent) e = cudaMemsetAsync(d_pc.p, 0, std::max<uint32_t>(n_parents, 1) * 8, s);
    if (e == cudaSuccess && group_parent) e = cudaMemcpyAsync(d_gp.p, group_parent, n_groups * 4, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess)
        e = launch_count_distinct(d_keys, d_off, n_keys, d_gb, n_groups, group_parent ? d_gp.p : nullptr, n_parents, d_em.p,
                                  d_gc.p, group_parent ? d_pc.p : nullptr, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_group_counts, d_gc.p, n_groups * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && group_parent)
        e = cudaMemcpyAsync(out_parent_counts, d_pc.p, n_parents * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(BSG_ERR_CUDA, "count_distinct: %s", cudaGetErrorString(e));
    return BSG_OK;
}

// (f.3) exact distinct counts per group and per parent union — replaces the dedup the Go maps do.
extern "C" int bsg_count_distinct(bsg_ctx* ctx, const uint8_t* keys, const uint64_t* key_off, uint64_t n_keys,
                                  const uint64_t* group_begin, uint32_t n_groups, const uint32_t* group_parent,
                                  uint32_t n_parents, uint64_t* out_group_counts, uint64_t* out_parent_counts) {
    if (!ctx || !key_off || (n_groups && (!group_begin || !out_group_counts)) ||
        ((group_parent != nullptr) != (out_parent_counts != nullptr)))
        return fail(BSG_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t nbytes = n_keys ? key_off[n_keys] : 0;
    if (nbytes && !keys) return fail(BSG_ERR_INVALID, "keys is NULL");
    {
        const uint64_t bad = first_bad_offset(key_off, n_keys);
        if (bad != n_keys) return fail(BSG_ERR_INVALID, "key_off not monotone at %llu", (unsigned long long)bad);
    }
    for (uint32_t g = 0; g < n_groups; ++g) {
        if (group_begin[g + 1] < group_begin[g] || group_begin[g + 1] > n_keys)
            return fail(BSG_ERR_INVALID, "group %u: key range invalid", g);
        if (group_parent && group_parent[g] >= n_parents) return fail(BSG_ERR_INVALID, "group %u: parent out of range", g);
    }
    if (n_groups && (group_begin[0] != 0 || group_begin[n_groups] != n_keys))
        return fail(BSG_ERR_INVALID, "groups must cover all keys");
    for (uint32_t g = 0; g < n_groups; ++g) out_group_counts[g] = 0;
    for (uint32_t p = 0; group_parent && p < n_parents; ++p) out_parent_counts[p] = 0;
    if (n_keys == 0 || n_groups == 0) return BSG_OK;
    cudaStream_t s = pool_get(ctx);
    if (!s) return fail(BSG_ERR_CUDA, "stream create failed");
    DevBuf<uint8_t> d_keys;
    DevBuf<uint64_t> d_off, d_gb;
    int rc = BSG_OK;
    do {
        if (d_keys.alloc(nbytes + kKeyPad) != cudaSuccess || d_off.alloc(n_keys + 1) != cudaSuccess ||
            d_gb.alloc(n_groups + 1) != cudaSuccess) {
            rc = fail(BSG_ERR_NOMEM, "device alloc");
            break;
        }
        cudaError_t e;
        {
            std::unique_lock<std::mutex> stage_lk(ctx->stage_mu);
            e = cudaMemsetAsync(d_keys.p + nbytes, 0, kKeyPad, s);
            if (e == cudaSuccess) e = upload(ctx, d_keys.p, keys, nbytes, s);
            if (e == cudaSuccess) e = upload(ctx, d_off.p, key_off, (n_keys + 1) * 8, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_gb.p, group_begin, (static_cast<size_t>(n_groups) + 1) * 8, cudaMemcpyHostToDevice, s);
        }
        if (e != cudaSuccess) { rc = fail(BSG_ERR_CUDA, "bsg_count_distinct: %s", cudaGetErrorString(e)); break; }
        rc = count_distinct_device(ctx, d_keys.p, d_off.p, n_keys, d_gb.p, n_groups, group_parent, n_parents, out_group_counts,
                                   out_parent_counts, s);
    } while (0);
    pool_put(ctx, s);
    return rc;
}

// ---------------------------------------------------------------- key sets ---
// A key set is a batch of grouped keys resident in HBM: the build side's counterpart of bsg_query.  The
// emissions of a flush cross PCIe ONCE; exact distinct counts (bsg_keyset_count_distinct), filter sizing
// on the host, and the build (bsg_keyset_build) then all read the resident copy.  bsg_keyset_build is
// asynchronous on the ctx stream and can write into any device buffer, e.g. symmetric memory that
// bsg_or_reduce_device combines across GPUs afterwards.
struct bsg_keyset {
    int device = 0;
    uint64_t n_keys = 0;
    uint32_t n_groups = 0;
    uint8_t* d_keys = nullptr;
    uint64_t* d_off = nullptr;
    uint64_t* d_gb = nullptr;            // caller's group_begin
    std::vector<uint64_t> h_gb;
    // set by bsg_keyset_set_filters
    bool planned = false;
    uint32_t n_sub = 0, smem_cap = 0, n_filters = 0;
    bool has_gf2 = false;
    uint64_t n_words = 0;
    uint64_t* d_sub_gb = nullptr;
    uint32_t *d_gf = nullptr, *d_gf2 = nullptr;
    BuildFilter* d_bf = nullptr;
    uint64_t* d_out = nullptr;           // internal output (when the caller passes no buffer)
    size_t cap_out = 0;
    const uint64_t* last_out = nullptr;  // where the last build wrote
};

extern "C" void bsg_keyset_free(bsg_keyset* k) {
    if (!k) return;
    cudaSetDevice(k->device);
    cudaFree(k->d_keys);
    cudaFree(k->d_off);
    cudaFree(k->d_gb);
    cudaFree(k->d_sub_gb);
    cudaFree(k->d_gf);
    cudaFree(k->d_gf2);
    cudaFree(k->d_bf);
    cudaFree(k->d_out);
    delete k;
}

extern "C" int bsg_keyset_create(bsg_ctx* ctx, const uint8_t* keys, const uint64_t* key_off, uint64_t n_keys,
                                 const uint64_t* group_begin, uint32_t n_groups, bsg_keyset** out) {
    if (!ctx || !out || !key_off || (n_groups && !group_begin)) return fail(BSG_ERR_INVALID, "NULL argument");
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t nbytes = n_keys ? key_off[n_keys] : 0;
    if (nbytes && !keys) return fail(BSG_ERR_INVALID, "keys is NULL");
    {
        const uint64_t bad = first_bad_offset(key_off, n_keys);
        if (bad != n_keys) return fail(BSG_ERR_INVALID, "key_off not monotone at %llu", (unsigned long long)bad);
    }
    for (uint32_t g = 0; g < n_groups; ++g)
        if (group_begin[g + 1] < group_begin[g] || group_begin[g + 1] > n_keys)
            return fail(BSG_ERR_INVALID, "group %u: key range invalid", g);
    if (n_groups && (group_begin[0] != 0 || group_begin[n_groups] != n_keys))
        return fail(BSG_ERR_INVALID, "groups must cover all keys");
    bsg_keyset* k = new (std::nothrow) bsg_keyset();
    if (!k) return fail(BSG_ERR_NOMEM, "keyset alloc");
    k->device = ctx->device;
    k->n_keys = n_keys;
    k->n_groups = n_groups;
    k->h_gb.assign(group_begin, group_begin + (n_groups ? n_groups + 1 : 0));
    if (k->h_gb.empty()) k->h_gb.push_back(0);
    cudaStream_t s = pool_get(ctx);
    if (!s) { delete k; return fail(BSG_ERR_CUDA, "stream create failed"); }
    auto body = [&]() -> int {
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&k->d_keys), nbytes + kKeyPad));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&k->d_off), (n_keys + 1) * 8));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&k->d_gb), k->h_gb.size() * 8));
        std::unique_lock<std::mutex> stage_lk(ctx->stage_mu);
        CUDA_TRY(cudaMemsetAsync(k->d_keys + nbytes, 0, kKeyPad, s));
        CUDA_TRY(upload(ctx, k->d_keys, keys, nbytes, s));
        CUDA_TRY(upload(ctx, k->d_off, key_off, (n_keys + 1) * 8, s));
        CUDA_TRY(cudaMemcpyAsync(k->d_gb, k->h_gb.data(), k->h_gb.size() * 8, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        return BSG_OK;
    };
    int rc = body();
    pool_put(ctx, s);
    if (rc) { bsg_keyset_free(k); return rc; }
    *out = k;
    return BSG_OK;
}

extern "C" int bsg_keyset_count_distinct(bsg_ctx* ctx, bsg_keyset* k, const uint32_t* group_parent, uint32_t n_parents,
                                         uint64_t* out_group_counts, uint64_t* out_parent_counts) {
    if (!ctx || !k || (k->n_groups && !out_group_counts) || ((group_parent != nullptr) != (out_parent_counts != nullptr)))
        return fail(BSG_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint32_t n_groups = k->n_groups;
    for (uint32_t g = 0; g < n_groups; ++g) {
        out_group_counts[g] = 0;
        if (group_parent && group_parent[g] >= n_parents) return fail(BSG_ERR_INVALID, "group %u: parent out of range", g);
    }
    for (uint32_t p = 0; group_parent && p < n_parents; ++p) out_parent_counts[p] = 0;
    if (k->n_keys == 0 || n_groups == 0) return BSG_OK;
    cudaStream_t s = pool_get(ctx);
    if (!s) return fail(BSG_ERR_CUDA, "stream create failed");
    int rc = count_distinct_device(ctx, k->d_keys, k->d_off, k->n_keys, k->d_gb, n_groups, group_parent, n_parents,
                                   out_group_counts, out_parent_counts, s);
    pool_put(ctx, s);
    return rc;
}

extern "C" int bsg_keyset_set_filters(bsg_ctx* ctx, bsg_keyset* k, const uint32_t* group_filter, const uint32_t* group_filter2,
                                      const bsg_filter_desc* desc, uint32_t n_filters, uint64_t n_words) {
    if (!ctx || !k || !desc || (k->n_groups && !group_filter)) return fail(BSG_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    BuildPlan P;
    int rc = plan_build(ctx, k->n_keys, k->h_gb.data(), k->n_groups, group_filter, group_filter2, desc, n_filters, n_words, P);
    if (rc) return rc;
    cudaFree(k->d_sub_gb); cudaFree(k->d_gf); cudaFree(k->d_gf2); cudaFree(k->d_bf);
    k->d_sub_gb = nullptr; k->d_gf = nullptr; k->d_gf2 = nullptr; k->d_bf = nullptr;
    k->planned = false;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&k->d_sub_gb), P.gbe.size() * 8));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&k->d_gf), std::max<size_t>(P.gf.size(), 1) * 4));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&k->d_bf), std::max<size_t>(P.bf.size(), 1) * sizeof(BuildFilter)));
    CUDA_TRY(cudaMemcpy(k->d_sub_gb, P.gbe.data(), P.gbe.size() * 8, cudaMemcpyHostToDevice));
    if (!P.gf.empty()) CUDA_TRY(cudaMemcpy(k->d_gf, P.gf.data(), P.gf.size() * 4, cudaMemcpyHostToDevice));
    if (!P.bf.empty()) CUDA_TRY(cudaMemcpy(k->d_bf, P.bf.data(), P.bf.size() * sizeof(BuildFilter), cudaMemcpyHostToDevice));
    k->has_gf2 = group_filter2 != nullptr;
    if (k->has_gf2) {
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&k->d_gf2), std::max<size_t>(P.gf2.size(), 1) * 4));
        if (!P.gf2.empty()) CUDA_TRY(cudaMemcpy(k->d_gf2, P.gf2.data(), P.gf2.size() * 4, cudaMemcpyHostToDevice));
    }
    k->n_sub = P.n_sub;
    k->smem_cap = P.smem_cap;
    k->n_filters = n_filters;
    k->n_words = n_words;
    k->planned = true;
    return BSG_OK;
}

extern "C" int bsg_keyset_build(bsg_ctx* ctx, bsg_keyset* k, uint64_t* d_out_words) {
    if (!ctx || !k) return fail(BSG_ERR_INVALID, "NULL argument");
    if (!k->planned) return fail(BSG_ERR_INVALID, "bsg_keyset_set_filters has not been called");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->cur_stream;
    uint64_t* out = d_out_words;
    if (!out) {
        if (k->cap_out < k->n_words * 8 || !k->d_out) {
            cudaFree(k->d_out);
            k->d_out = nullptr;
            k->cap_out = 0;
            CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&k->d_out), std::max<uint64_t>(k->n_words, 1) * 8));
            k->cap_out = std::max<uint64_t>(k->n_words, 1) * 8;
        }
        out = k->d_out;
    }
    CUDA_TRY(cudaMemsetAsync(out, 0, std::max<uint64_t>(k->n_words, 1) * 8, s));
    CUDA_TRY(launch_build(k->d_keys, k->d_off, k->d_sub_gb, k->n_sub, k->d_gf, k->has_gf2 ? k->d_gf2 : nullptr, k->d_bf, out,
                          k->smem_cap, s));
    k->last_out = out;
    return BSG_OK;
}

extern "C" int bsg_keyset_fetch(bsg_ctx* ctx, bsg_keyset* k, uint64_t* out_words) {
    if (!ctx || !k || !out_words) return fail(BSG_ERR_INVALID, "NULL argument");
    if (!k->last_out) return fail(BSG_ERR_INVALID, "bsg_keyset_build has not been called");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->cur_stream));
    if (k->n_words == 0) return BSG_OK;
    if (k->n_words * 8 < kStageThreshold) {
        CUDA_TRY(cudaMemcpy(out_words, k->last_out, k->n_words * 8, cudaMemcpyDeviceToHost));
    } else {
        std::unique_lock<std::mutex> stage_lk(ctx->stage_mu);
        CUDA_TRY(staged_copy(ctx, const_cast<uint64_t*>(k->last_out), out_words, k->n_words * 8, false));
    }
    return BSG_OK;
}

extern "C" const uint64_t* bsg_keyset_device_words(const bsg_keyset* k) { return k ? k->last_out : nullptr; }

// ------------------------------------------------------------------ corpus ---
struct bsg_corpus {
    int device = 0;
    uint64_t n_units = 0;
    DevFilter* d_udesc = nullptr;
    StageRow* d_stab = nullptr;  // one row per staged unit, staged order
    uint64_t* d_words = nullptr;
    uint64_t total_words = 0;
    uint32_t* d_staged_list = nullptr;  // nullptr when every unit is staged (identity list)
    uint32_t n_staged = 0;
    uint32_t* d_gather_list = nullptr;
    uint32_t n_gather = 0;
    uint32_t* d_bad32 = nullptr;   // sections loader: bit u = unit u's filter section failed to parse (never a candidate)
    uint32_t* d_parent = nullptr;  // optional: unit -> index of its parent unit in another corpus (block -> file)
    uint64_t n_parents = 0;        // number of units in the parent corpus
    uint32_t stage_cap_bytes = 0;  // largest staged unit (all kinds), 16-byte multiple
    uint64_t kind_bytes[3] = {0, 0, 0};         // Σ 8*ceil(m/64) per kind
    uint64_t staged_kind_bytes[3] = {0, 0, 0};  // same, staged units only (padded words)
    std::vector<bsg_filter_desc> h_desc;        // caller's (m,k) with word_off in the device layout
    // tile ring (probe_tiles_kernel): records, which units they cover, and the rest (gather list)
    TileRec* d_tiles = nullptr;
    uint32_t t_items = 0;          // items = tiles (UNIT mode) or staged units (KIND mode, 2 tiles each)
    uint32_t t_parts = 1;
    uint32_t t_units_cap = 1;      // most units in one tile (stage header size)
    uint32_t t_data_cap = 0;       // largest tile data, 128-byte multiple
    uint32_t t_staged = 0;         // units covered by tiles
    uint32_t* d_t_staged_list = nullptr;  // nullptr when every unit is tiled
    uint32_t* d_t_gather_list = nullptr;
    uint32_t t_gather = 0;
    uint64_t t_staged_kind_bytes[3] = {0, 0, 0};
};

extern "C" void bsg_corpus_free(bsg_corpus* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_udesc);
    cudaFree(c->d_stab);
    cudaFree(c->d_words);
    cudaFree(c->d_staged_list);
    cudaFree(c->d_gather_list);
    cudaFree(c->d_parent);
    cudaFree(c->d_bad32);
    cudaFree(c->d_tiles);
    cudaFree(c->d_t_staged_list);
    cudaFree(c->d_t_gather_list);
    delete c;
}

extern "C" uint64_t bsg_corpus_units(const bsg_corpus* c) { return c ? c->n_units : 0; }

extern "C" uint64_t bsg_corpus_bitset_bytes(const bsg_corpus* c, uint32_t kind_mask) {
    if (!c) return 0;
    uint64_t t = 0;
    for (int k = 0; k < 3; ++k)
        if (kind_mask & (1u << k)) t += c->kind_bytes[k];
    return t;
}

// HBM held by a corpus: bitset words + descriptor / stage / tile tables + unit lists (the resident cache's unit of account)
extern "C" uint64_t bsg_corpus_device_bytes(const bsg_corpus* c) {
    if (!c) return 0;
    uint64_t b = (c->total_words + 2) * 8 + c->n_units * 3 * sizeof(DevFilter);
    b += static_cast<uint64_t>(c->n_staged) * sizeof(StageRow) + static_cast<uint64_t>(c->t_items) * c->t_parts * sizeof(TileRec);
    b += 4ull * (c->d_staged_list ? c->n_staged : 0) + 4ull * c->n_gather + 4ull * (c->d_t_staged_list ? c->t_staged : 0) + 4ull * c->t_gather;
    if (c->d_parent) b += c->n_units * 4;
    if (c->d_bad32) b += (c->n_units + 31) / 32 * 4 + 8;
    return b;
}

extern "C" int bsg_corpus_unit_desc(const bsg_corpus* c, uint64_t unit, bsg_filter_desc out_desc[3]) {
    if (!c || !out_desc || unit >= c->n_units) return fail(BSG_ERR_INVALID, "bad unit");
    for (int k = 0; k < 3; ++k) out_desc[k] = c->h_desc[unit * 3 + k];
    return BSG_OK;
}

namespace {

// Shared tail of the two corpus loaders: given per-slot (m,k) [m==0 absent], lays
// the corpus out, uploads descriptor tables and decides staged vs gather units.
struct Layout {
    std::vector<DevFilter> udesc;
    std::vector<UnitTab> utab;
    uint64_t total_words = 0;
};

int make_layout(const bsg_filter_desc* desc, uint64_t n_units, Layout& L) {
    L.udesc.resize(n_units * 3);
    L.utab.resize(n_units);
    uint64_t cur = 0;
    for (uint64_t u = 0; u < n_units; ++u) {
        UnitTab& t = L.utab[u];
        t.word_base = cur;
        t.total = 0;
        t.reserved = 0;
        uint64_t unit_words = 0;
        for (int k = 0; k < 3; ++k) {
            const bsg_filter_desc& d = desc[u * 3 + k];
            DevFilter& f = L.udesc[u * 3 + k];
            if (d.m == 0) { f = DevFilter{cur + unit_words, 0, 0, 0, 0}; t.nw[k] = 0; continue; }
            const uint64_t nw = words_for(d.m);
            f.word_off = cur + unit_words;
            f.m = d.m;
            f.inv = reciprocal(d.m);
            f.k = static_cast<uint32_t>(d.k);
            f.nwords = static_cast<uint32_t>(nw);
            const uint64_t padded = even_up(nw);
            if (unit_words + padded > 0xffffffffull) return fail(BSG_ERR_INVALID, "unit %llu too large", (unsigned long long)u);
            t.nw[k] = static_cast<uint32_t>(padded);
            unit_words += padded;
        }
        t.total = static_cast<uint32_t>(unit_words);
        cur += unit_words;
    }
    L.total_words = cur;
    return BSG_OK;
}

// Shared memory one CTA of the configured probe_tiles shape may use: 1024 / threads CTAs share an SM, each
// pays 1 KB of system-reserved shared memory.
uint64_t tiles_cta_smem(const bsg_ctx* ctx) {
    const uint64_t per_sm = 1024 / static_cast<uint64_t>(probe_tiles_threads(ctx->tiles_shape));
    const uint64_t sm_total = static_cast<uint64_t>(ctx->max_smem_optin) + 1024;   // 228 KB per SM, 227 KB opt-in per CTA
    return std::min<uint64_t>(ctx->max_smem_optin, (sm_total / per_sm - 1024) & ~uint64_t(127));
}

// Cuts the corpus into tiles for probe_tiles_kernel (bsg_internal.h): decides UNIT vs KIND mode, groups
// small units, and lists the units no tile can hold (they take the gather kernel).
int build_tiles(bsg_ctx* ctx, bsg_corpus* c, const Layout& L, cudaStream_t s) {
    const uint64_t n_units = c->n_units;
    // ring budget for a full pass of keys: the fixed part grows with the units per tile (survivor lists, rows).
    const uint32_t group_cap_cfg = static_cast<uint32_t>(std::min<int>(ctx->tile_units, kTileMaxUnits));
    const uint64_t cta_smem = tiles_cta_smem(ctx);
    auto ring_budget = [&](uint32_t units_cap) {
        return cta_smem - tiles_fixed_smem(units_cap, kProbeMaxKeysPerPass);
    };
    // several small units per tile: two stages are enough (the kernel is instruction bound there and a larger tile
    // amortises the per-tile rounds: 2a 53.2 -> 50.7 us); a tile that is one large unit keeps a ring of >= 3
    auto tile_limit = [&](uint32_t g) {
        const uint64_t min_stages = ctx->tile_min_stages > 0 ? static_cast<uint64_t>(ctx->tile_min_stages) : (g > 1 ? 2 : 3);
        return ring_budget(g) / min_stages - tile_header_bytes(g);
    };
    // units per tile: the g in 1..cfg that packs the most units of the typical size (more units per tile cost list space)
    uint32_t group_cap = 1;
    {
        uint64_t sum = 0, cnt = 0;
        for (uint64_t u = 0; u < n_units; ++u) {
            const uint64_t b = static_cast<uint64_t>(L.utab[u].total) * 8;
            if (b && b <= tile_limit(1)) { sum += b; ++cnt; }
        }
        const uint64_t typical = cnt ? std::max<uint64_t>(sum / cnt, 16) : 16;
        uint64_t best = 0;
        for (uint32_t g = 1; g <= group_cap_cfg; ++g) {
            const uint64_t lim = std::min<uint64_t>(tile_limit(g), static_cast<uint64_t>(ctx->tile_bytes));
            const uint64_t fit = std::min<uint64_t>(g, lim / typical);
            if (fit > best) { best = fit; group_cap = g; }
        }
    }
    const uint64_t unit_limit = tile_limit(group_cap);                         // UNIT mode: largest tile (and unit)
    const uint64_t part_limit = ring_budget(1) / 2 - tile_header_bytes(1);     // KIND mode: >= 2 stages (lock-step kernel)
    auto part_bytes = [&](uint64_t u, int part) -> uint64_t {
        const UnitTab& t = L.utab[u];
        return part == 0 ? (static_cast<uint64_t>(t.nw[0]) + t.nw[1]) * 8 : static_cast<uint64_t>(t.nw[2]) * 8;
    };
    auto filters_ok = [&](uint64_t u) {
        for (int k = 0; k < 3; ++k) {
            const DevFilter& d = L.udesc[u * 3 + k];
            if (d.m && (d.m >= (1ull << 30) || d.k > kTileMaxK)) return false;
        }
        return true;
    };
    uint64_t bytes_unit = 0, bytes_kind_only = 0;
    std::vector<uint8_t> cls(n_units, 0);  // 1 = fits UNIT mode, 2 = fits KIND mode only, 0 = neither
    for (uint64_t u = 0; u < n_units; ++u) {
        if (!filters_ok(u)) continue;
        const uint64_t b = static_cast<uint64_t>(L.utab[u].total) * 8;
        if (b <= unit_limit) { cls[u] = 1; bytes_unit += b; }
        else if (std::max(part_bytes(u, 0), part_bytes(u, 1)) <= part_limit) { cls[u] = 2; bytes_kind_only += b; }
    }
    bool kind_mode = bytes_kind_only * 10 > bytes_unit + bytes_kind_only;
    if (ctx->tile_mode == 1) kind_mode = false;
    if (ctx->tile_mode == 2) kind_mode = true;
    std::vector<TileRec> recs;
    std::vector<uint32_t> staged, gather;
    uint32_t units_cap = 1;
    uint64_t data_cap = 0;
    auto tile_filter = [&](uint64_t u, int k, uint64_t rel_bytes) {
        const DevFilter& d = L.udesc[u * 3 + k];
        if (d.m == 0) return tile_filter_absent();
        // offsets count from the start of the stage (the kernel adds one base, not header + data)
        return TileFilter{static_cast<uint32_t>(d.m), static_cast<uint32_t>(d.inv >> 32), static_cast<uint32_t>(d.inv),
                          (static_cast<uint32_t>(rel_bytes + kTileBitmapOff) << 8) | static_cast<uint32_t>(d.k)};
    };
    auto small_k = [&](uint64_t u, int k) { const DevFilter& d = L.udesc[u * 3 + k]; return d.m && d.k < 4; };
    if (kind_mode) {
        for (uint64_t u = 0; u < n_units; ++u) {
            if (cls[u] == 0) { gather.push_back(static_cast<uint32_t>(u)); continue; }
            staged.push_back(static_cast<uint32_t>(u));
            const UnitTab& t = L.utab[u];
            for (int part = 0; part < 2; ++part) {
                TileRec r;
                memset(&r, 0, sizeof(r));
                r.ones = 0xffffffffu;
                for (auto& fu : r.f) for (auto& fk : fu) fk = tile_filter_absent();
                r.n_units = 1;
                r.part_kinds = part == 0 ? 3u : 4u;
                r.flags = part == 0 ? kTileFirstPart : kTileLastPart;
                r.unit[0] = static_cast<uint32_t>(u);
                r.fill.rec_bytes = kTileRecFixedBytes + 48;
                if (part == 0) {
                    r.fill.word_base = t.word_base;
                    r.fill.data_bytes = (t.nw[0] + t.nw[1]) * 8u;
                    r.fill.nb16[0][0] = static_cast<uint16_t>(t.nw[0] / 2);
                    r.fill.nb16[0][1] = static_cast<uint16_t>(t.nw[1] / 2);
                    r.f[0][0] = tile_filter(u, 0, 0);
                    r.f[0][1] = tile_filter(u, 1, static_cast<uint64_t>(t.nw[0]) * 8);
                    if (small_k(u, 0) || small_k(u, 1)) r.flags |= kTileSmallK;
                } else {
                    r.fill.word_base = t.word_base + t.nw[0] + t.nw[1];
                    r.fill.data_bytes = t.nw[2] * 8u;
                    r.fill.nb16[0][2] = static_cast<uint16_t>(t.nw[2] / 2);
                    r.f[0][2] = tile_filter(u, 2, 0);
                    if (small_k(u, 2)) r.flags |= kTileSmallK;
                }
                data_cap = std::max<uint64_t>(data_cap, r.fill.data_bytes);
                recs.push_back(r);
            }
        }
    } else {
        const uint32_t group_max = group_cap;
        const uint64_t target = static_cast<uint64_t>(ctx->tile_bytes);
        TileRec r;
        memset(&r, 0, sizeof(r));
        uint64_t cur_bytes = 0, next_word = 0;
        bool open = false;
        auto close_tile = [&]() {
            if (!open) return;
            r.fill.rec_bytes = kTileRecFixedBytes + 48 * r.n_units;
            r.fill.data_bytes = static_cast<uint32_t>(cur_bytes);
            units_cap = std::max(units_cap, r.n_units);
            data_cap = std::max(data_cap, cur_bytes);
            recs.push_back(r);
            open = false;
        };
        for (uint64_t u = 0; u < n_units; ++u) {
            if (cls[u] != 1) { gather.push_back(static_cast<uint32_t>(u)); continue; }
            staged.push_back(static_cast<uint32_t>(u));
            const UnitTab& t = L.utab[u];
            const uint64_t b = static_cast<uint64_t>(t.total) * 8;
            // a tile is one contiguous range of the words array: close it at gaps (gather units in between)
            if (open && (r.n_units >= group_max || cur_bytes + b > std::min(target, unit_limit) || t.word_base != next_word))
                close_tile();
            if (!open) {
                memset(&r, 0, sizeof(r));
                r.ones = 0xffffffffu;
                for (auto& fu : r.f) for (auto& fk : fu) fk = tile_filter_absent();
                r.part_kinds = 7u;
                r.flags = kTileFirstPart | kTileLastPart;
                r.fill.word_base = t.word_base;
                cur_bytes = 0;
                open = true;
            }
            const uint32_t j = r.n_units++;
            r.unit[j] = static_cast<uint32_t>(u);
            uint64_t rel = cur_bytes;
            for (int k = 0; k < 3; ++k) {
                r.f[j][k] = tile_filter(u, k, rel);
                r.fill.nb16[j][k] = static_cast<uint16_t>(t.nw[k] / 2);
                rel += static_cast<uint64_t>(t.nw[k]) * 8;
                if (small_k(u, k)) r.flags |= kTileSmallK;
            }
            cur_bytes += b;
            next_word = t.word_base + t.total;
        }
        close_tile();
    }
    c->t_parts = kind_mode ? 2u : 1u;
    c->t_items = static_cast<uint32_t>(recs.size() / c->t_parts);
    c->t_units_cap = units_cap;
    c->t_data_cap = static_cast<uint32_t>((std::max<uint64_t>(data_cap, 16) + 127) & ~127ull);
    c->t_staged = static_cast<uint32_t>(staged.size());
    c->t_gather = static_cast<uint32_t>(gather.size());
    for (uint32_t u : staged)
        for (int k = 0; k < 3; ++k)
            if (L.udesc[static_cast<uint64_t>(u) * 3 + k].m) c->t_staged_kind_bytes[k] += static_cast<uint64_t>(L.utab[u].nw[k]) * 8;
    if (!recs.empty()) {
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_tiles), recs.size() * sizeof(TileRec)));
        CUDA_TRY(cudaMemcpyAsync(c->d_tiles, recs.data(), recs.size() * sizeof(TileRec), cudaMemcpyHostToDevice, s));
    }
    if (!gather.empty()) {
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_t_gather_list), gather.size() * 4));
        CUDA_TRY(cudaMemcpyAsync(c->d_t_gather_list, gather.data(), gather.size() * 4, cudaMemcpyHostToDevice, s));
        if (!staged.empty()) {
            CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_t_staged_list), staged.size() * 4));
            CUDA_TRY(cudaMemcpyAsync(c->d_t_staged_list, staged.data(), staged.size() * 4, cudaMemcpyHostToDevice, s));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(s));  // the host vectors die with this scope
    return BSG_OK;
}

int finish_corpus(bsg_ctx* ctx, bsg_corpus* c, const Layout& L, const bsg_filter_desc* desc, cudaStream_t s) {
    const uint64_t n_units = c->n_units;
    // staged vs gather: a unit is staged when at least 3 stages of its size fit
    // (sized for the two-phase kernel's larger stage header so either staged kernel can run)
    const uint64_t budget = static_cast<uint64_t>(ctx->max_smem_optin) - kProbe2SmemPrefixBytes;
    const uint64_t unit_limit = budget / 3 - kProbeStage2HeaderBytes;
    std::vector<uint32_t> staged, gather;
    std::vector<StageRow> stab;
    uint32_t cap = 0;
    c->h_desc.resize(n_units * 3);
    for (uint64_t u = 0; u < n_units; ++u) {
        const uint64_t bytes = static_cast<uint64_t>(L.utab[u].total) * 8;
        const bool st = bytes <= unit_limit;
        if (st) {
            staged.push_back(static_cast<uint32_t>(u));
            cap = std::max<uint32_t>(cap, static_cast<uint32_t>(bytes));
            StageRow r;
            r.unit = static_cast<uint32_t>(u);
            r.total_words = L.utab[u].total;
            r.word_base = L.utab[u].word_base;
            uint32_t rel = 0;
            for (int k = 0; k < 3; ++k) {
                const DevFilter& d = L.udesc[u * 3 + k];
                r.nw[k] = L.utab[u].nw[k];
                r.f[k] = StageFilter{static_cast<uint32_t>(d.m), d.k, static_cast<uint32_t>(d.inv >> 32),
                                     static_cast<uint32_t>(d.inv), rel, {0, 0, 0}};
                rel += L.utab[u].nw[k] * 8u;
            }
            r.pad = 0;
            stab.push_back(r);
        } else {
            gather.push_back(static_cast<uint32_t>(u));
        }
        for (int k = 0; k < 3; ++k) {
            const bsg_filter_desc& d = desc[u * 3 + k];
            c->h_desc[u * 3 + k] = bsg_filter_desc{d.m, d.m ? d.k : 0, L.udesc[u * 3 + k].word_off};
            if (d.m) {
                c->kind_bytes[k] += words_for(d.m) * 8;
                if (st) c->staged_kind_bytes[k] += static_cast<uint64_t>(L.utab[u].nw[k]) * 8;
            }
        }
    }
    c->stage_cap_bytes = (cap + 15u) & ~15u;
    c->n_staged = static_cast<uint32_t>(staged.size());
    c->n_gather = static_cast<uint32_t>(gather.size());
    if (!stab.empty()) {
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_stab), stab.size() * sizeof(StageRow)));
        CUDA_TRY(cudaMemcpyAsync(c->d_stab, stab.data(), stab.size() * sizeof(StageRow), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaStreamSynchronize(s));  // stab dies with this scope
    }
    if (!gather.empty()) {
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_gather_list), gather.size() * 4));
        CUDA_TRY(cudaMemcpyAsync(c->d_gather_list, gather.data(), gather.size() * 4, cudaMemcpyHostToDevice, s));
        if (!staged.empty()) {
            CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_staged_list), staged.size() * 4));
            CUDA_TRY(cudaMemcpyAsync(c->d_staged_list, staged.data(), staged.size() * 4, cudaMemcpyHostToDevice, s));
        }
        CUDA_TRY(cudaStreamSynchronize(s));  // vectors die with this scope
    }
    return build_tiles(ctx, c, L, s);
}

}  // namespace

extern "C" int bsg_corpus_load(bsg_ctx* ctx, const bsg_filter_desc* desc, uint64_t n_units, const uint64_t* words,
                               uint64_t n_words, int big_endian, bsg_corpus** out) {
    if (!ctx || !out || (n_units && !desc) || (n_words && !words)) return fail(BSG_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (n_units > 0xfffffff0ull / 3) return fail(BSG_ERR_INVALID, "too many units");
    CUDA_TRY(cudaSetDevice(ctx->device));
    for (uint64_t f = 0; f < n_units * 3; ++f) {
        int rc = validate_filter(desc[f], n_words, "filter slot", f);
        if (rc) return rc;
    }
    Layout L;
    int rc = make_layout(desc, n_units, L);
    if (rc) return rc;
    std::vector<uint64_t> src_off(n_units * 3);
    for (uint64_t f = 0; f < n_units * 3; ++f) src_off[f] = desc[f].word_off;

    bsg_corpus* c = new (std::nothrow) bsg_corpus();
    if (!c) return fail(BSG_ERR_NOMEM, "corpus alloc");
    c->device = ctx->device;
    c->n_units = n_units;
    c->total_words = L.total_words;
    cudaStream_t s = pool_get(ctx);
    if (!s) { delete c; return fail(BSG_ERR_CUDA, "stream create failed"); }
    DevBuf<uint64_t> d_src, d_src_off;
    auto body = [&]() -> int {
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_udesc), std::max<uint64_t>(n_units * 3, 1) * sizeof(DevFilter)));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_words), (L.total_words + 2) * 8));
        CUDA_TRY(d_src.alloc(n_words));
        CUDA_TRY(d_src_off.alloc(n_units * 3));
        CUDA_TRY(cudaMemsetAsync(c->d_words, 0, (L.total_words + 2) * 8, s));
        if (n_units) {
            CUDA_TRY(cudaMemcpyAsync(c->d_udesc, L.udesc.data(), n_units * 3 * sizeof(DevFilter), cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(d_src_off.p, src_off.data(), n_units * 3 * 8, cudaMemcpyHostToDevice, s));
        }
        if (n_words) {
            std::unique_lock<std::mutex> stage_lk(ctx->stage_mu);
            CUDA_TRY(upload(ctx, d_src.p, words, n_words * 8, s));
        }
        CUDA_TRY(launch_repack(d_src.p, d_src_off.p, c->d_udesc, n_units * 3, c->d_words, big_endian, s));
        int r = finish_corpus(ctx, c, L, desc, s);
        if (r) return r;
        CUDA_TRY(cudaStreamSynchronize(s));
        return BSG_OK;
    };
    rc = body();
    pool_put(ctx, s);
    if (rc) { bsg_corpus_free(c); return rc; }
    *out = c;
    return BSG_OK;
}

extern "C" int bsg_corpus_load_sections(bsg_ctx* ctx, const uint8_t* sections, const uint64_t* sec_off,
                                        uint64_t n_units, int verify_crc, int32_t* unit_status, uint64_t* n_bad,
                                        bsg_corpus** out) {
    if (!ctx || !out || (n_units && (!sec_off || !sections))) return fail(BSG_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (n_bad) *n_bad = 0;
    if (n_units > 0xfffffff0ull / 3) return fail(BSG_ERR_INVALID, "too many units");
    CUDA_TRY(cudaSetDevice(ctx->device));
    for (uint64_t u = 0; u < n_units; ++u)
        if (sec_off[u + 1] < sec_off[u]) return fail(BSG_ERR_INVALID, "sec_off not monotone at %llu", (unsigned long long)u);
    const uint64_t total = n_units ? sec_off[n_units] : 0;

    bsg_corpus* c = new (std::nothrow) bsg_corpus();
    if (!c) return fail(BSG_ERR_NOMEM, "corpus alloc");
    c->device = ctx->device;
    c->n_units = n_units;
    cudaStream_t s = pool_get(ctx);
    if (!s) { delete c; return fail(BSG_ERR_CUDA, "stream create failed"); }
    DevBuf<uint8_t> d_sec;
    DevBuf<uint64_t> d_off;
    DevBuf<SectionInfo> d_info;
    std::vector<SectionInfo> info(n_units);
    std::vector<bsg_filter_desc> desc(n_units * 3);
    uint64_t bad = 0;
    auto body = [&]() -> int {
        CUDA_TRY(d_sec.alloc(total + kKeyPad));
        CUDA_TRY(d_off.alloc(n_units + 1));
        CUDA_TRY(d_info.alloc(n_units));
        CUDA_TRY(cudaMemsetAsync(d_sec.p + total, 0, kKeyPad, s));
        if (total) {
            std::unique_lock<std::mutex> stage_lk(ctx->stage_mu);
            CUDA_TRY(upload(ctx, d_sec.p, sections, total, s));
        }
        if (n_units) CUDA_TRY(cudaMemcpyAsync(d_off.p, sec_off, (n_units + 1) * 8, cudaMemcpyHostToDevice, s));
        CUDA_TRY(launch_parse_sections(d_sec.p, d_off.p, n_units, verify_crc, d_info.p, s));
        if (n_units) CUDA_TRY(cudaMemcpyAsync(info.data(), d_info.p, n_units * sizeof(SectionInfo), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        std::vector<uint32_t> bad32((n_units + 31) / 32 + 2, 0u);
        for (uint64_t u = 0; u < n_units; ++u) {
            if (info[u].status != 0) { ++bad; bad32[u >> 5] |= 1u << (u & 31); }
            if (unit_status) unit_status[u] = info[u].status;
            for (int k = 0; k < 3; ++k) desc[u * 3 + k] = bsg_filter_desc{info[u].m[k], info[u].k[k], 0};
        }
        if (bad) {
            CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_bad32), bad32.size() * 4));
            CUDA_TRY(cudaMemcpy(c->d_bad32, bad32.data(), bad32.size() * 4, cudaMemcpyHostToDevice));
        }
        Layout L;
        int r = make_layout(desc.data(), n_units, L);
        if (r) return r;
        c->total_words = L.total_words;
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_udesc), std::max<uint64_t>(n_units * 3, 1) * sizeof(DevFilter)));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_words), (L.total_words + 2) * 8));
        CUDA_TRY(cudaMemsetAsync(c->d_words, 0, (L.total_words + 2) * 8, s));
        if (n_units)
            CUDA_TRY(cudaMemcpyAsync(c->d_udesc, L.udesc.data(), n_units * 3 * sizeof(DevFilter), cudaMemcpyHostToDevice, s));
        CUDA_TRY(launch_repack_sections(d_sec.p, d_info.p, c->d_udesc, n_units, c->d_words, s));
        r = finish_corpus(ctx, c, L, desc.data(), s);
        if (r) return r;
        CUDA_TRY(cudaStreamSynchronize(s));
        return BSG_OK;
    };
    int rc = body();
    pool_put(ctx, s);
    if (rc) { bsg_corpus_free(c); return rc; }
    if (n_bad) *n_bad = bad;
    *out = c;
    return BSG_OK;
}

// ------------------------------------------------------------------- query ---
struct bsg_query {
    int device = 0;
    // capacities (bytes) of the device buffers below: a query object can be re-prepared for
    // another batch without reallocating (bsg_probe keeps one per pooled stream)
    size_t cap_keys = 0, cap_key_off = 0, cap_kinds = 0, cap_hashes = 0, cap_prog = 0, cap_matrix = 0, cap_mask = 0;
    // pinned host staging (bsg_probe path): small inputs go up in async copies from here and the
    // results come down into it, then are memcpy'd to the caller's (pageable) buffers
    uint8_t* h_pin = nullptr;
    size_t cap_pin = 0;
    // device mirror of the pinned input block [keys + pad][offsets][kinds][program]: the bsg_probe
    // path uploads a batch with ONE async copy; k_* are the pointers the kernels read (they alias
    // d_in on that path, the separately owned buffers on the bsg_query_create path)
    uint8_t* d_in = nullptr;
    size_t cap_in = 0;
    const uint8_t* k_keys = nullptr;
    const uint64_t* k_key_off = nullptr;
    const uint8_t* k_kinds = nullptr;
    const bsg_expr_op* k_prog = nullptr;
    // probe_tiles: keys sorted by kind per 1024-key pass; slot -> caller index within the pass | kind << 14
    const uint16_t* k_slot = nullptr;
    uint16_t* d_slot = nullptr;
    size_t cap_slot = 0;
    std::vector<uint16_t> h_slot;
    TileRec* d_tiles_c = nullptr;  // hierarchical probes: the compacted tile records
    size_t cap_tiles_c = 0;
    // where the probe kernels write the (unit x key) matrix: d_matrix32, or — bsg_probe() without a
    // mask — the pinned host block h_out (zero copy: rows cross PCIe as posted writes while the
    // kernel runs, no D2H copy operation after it)
    // bsg_probe path: hashing is deferred to the run; a staged-only run hashes inside the probe kernel
    // (per-CTA scratch table), any other run launches hash_keys_kernel first
    bool hashed = false;
    uint64_t* d_hash_scratch = nullptr;
    size_t cap_hash_scratch = 0;
    uint32_t* k_matrix = nullptr;
    bool direct_out = false;       // k_matrix points into a caller buffer from bsg_host_alloc
    uint32_t* h_out = nullptr;
    uint32_t* h_out_dev = nullptr;
    size_t cap_out = 0;
    uint64_t out_units = ~0ull;
    uint32_t out_row_words32 = ~0u, out_groups = ~0u;
    // hierarchical probes: stage rows of the units whose parent survived + their count
    StageRow* d_rows = nullptr;
    size_t cap_rows = 0;
    uint32_t* d_n_rows = nullptr;
    // output shape the pad words were last zeroed for (kernels never write pad words)
    uint64_t zeroed_units = ~0ull;
    uint32_t zeroed_row_words32 = ~0u, zeroed_groups = ~0u;
    uint32_t n_keys = 0;
    uint32_t prog_len = 0;
    uint32_t kind_mask = 0;
    uint32_t row_words32 = 0;
    uint64_t n_units = 0;
    uint8_t* d_keys = nullptr;
    uint64_t* d_key_off = nullptr;
    uint8_t* d_kinds = nullptr;
    uint64_t* d_hashes = nullptr;
    bsg_expr_op* d_prog = nullptr;
    uint32_t* d_matrix32 = nullptr;
    uint32_t* d_mask32 = nullptr;
    uint32_t* d_multi_mask32 = nullptr;   // bsg_probe_multi: one candidate mask per query of the batch
    size_t cap_multi_mask = 0;
    int last_launches = 0;
};

extern "C" void bsg_query_free(bsg_query* q) {
    if (!q) return;
    cudaSetDevice(q->device);
    cudaFree(q->d_keys);
    cudaFree(q->d_key_off);
    cudaFree(q->d_kinds);
    cudaFree(q->d_hashes);
    cudaFree(q->d_prog);
    cudaFree(q->d_matrix32);
    cudaFree(q->d_mask32);
    cudaFree(q->d_multi_mask32);
    if (q->h_pin) cudaFreeHost(q->h_pin);
    if (q->h_out) cudaFreeHost(q->h_out);
    cudaFree(q->d_in);
    cudaFree(q->d_hash_scratch);
    cudaFree(q->d_rows);
    cudaFree(q->d_n_rows);
    cudaFree(q->d_slot);
    cudaFree(q->d_tiles_c);
    delete q;
}

extern "C" int bsg_query_last_launches(const bsg_query* q) { return q ? q->last_launches : 0; }

static int validate_program(const bsg_expr_op* prog, uint32_t prog_len, uint32_t n_keys) {
    uint32_t sp = 0;
    for (uint32_t pc = 0; pc < prog_len; ++pc) {
        const uint32_t op = prog[pc].op, arg = prog[pc].arg;
        switch (op) {
        case BSG_OP_LEAF:
            if (arg >= n_keys) return fail(BSG_ERR_INVALID, "program[%u]: leaf %u >= n_keys %u", pc, arg, n_keys);
            ++sp;
            break;
        case BSG_OP_TRUE: case BSG_OP_FALSE: ++sp; break;
        case BSG_OP_AND: case BSG_OP_OR:
            if (arg > sp) return fail(BSG_ERR_INVALID, "program[%u]: pops %u of %u", pc, arg, sp);
            sp = sp - arg + 1;
            break;
        default: return fail(BSG_ERR_INVALID, "program[%u]: unknown op %u", pc, op);
        }
        if (sp > BSG_MAX_STACK) return fail(BSG_ERR_INVALID, "program[%u]: stack deeper than %d", pc, BSG_MAX_STACK);
    }
    if (sp != 1) return fail(BSG_ERR_INVALID, "program leaves %u values on the stack (want 1)", sp);
    return BSG_OK;
}

template <typename T>
static cudaError_t ensure_cap(T*& p, size_t& cap, size_t need_bytes) {
    if (need_bytes <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = std::max<size_t>(need_bytes + need_bytes / 2, 256);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), want);
    if (e == cudaSuccess) cap = want;
    return e;
}

// Validates a batch and (re)fills query object q for it: uploads keys / kinds / program and
// launches the hash kernel on stream s.  Buffers only grow.
static int query_prepare_on(bsg_ctx* ctx, const bsg_corpus* corpus, const uint8_t* keys, const uint64_t* key_off,
                            uint32_t n_keys, const uint8_t* key_kind, const bsg_expr_op* prog, uint32_t prog_len,
                            cudaStream_t s, bsg_query* q, bool use_pinned = false, bool prog_prevalidated = false) {
    if (!ctx || !corpus || !q || (n_keys && (!key_off || !key_kind)) || (prog_len && !prog))
        return fail(BSG_ERR_INVALID, "NULL argument");
    const uint64_t nbytes = n_keys ? key_off[n_keys] : 0;
    if (nbytes && !keys) return fail(BSG_ERR_INVALID, "keys is NULL");
    uint32_t kind_mask = 0;
    for (uint32_t i = 0; i < n_keys; ++i) {
        if (key_off[i + 1] < key_off[i] || key_off[i + 1] - key_off[i] > 0xffffffffull)
            return fail(BSG_ERR_INVALID, "key_off not monotone at %u", i);
        if (key_kind[i] > 2) return fail(BSG_ERR_INVALID, "key %u: kind %u unknown", i, key_kind[i]);
        kind_mask |= 1u << key_kind[i];
    }
    if (prog_len && !prog_prevalidated) {
        int rc = validate_program(prog, prog_len, n_keys);
        if (rc) return rc;
    }
    // probe_tiles: stable counting sort of every 1024-key pass by kind
    q->h_slot.resize(std::max<uint32_t>(n_keys, 1));
    for (uint32_t kb = 0; kb < n_keys; kb += kProbeMaxKeysPerPass) {
        const uint32_t nk = std::min<uint32_t>(kProbeMaxKeysPerPass, n_keys - kb);
        uint32_t cnt[3] = {0, 0, 0};
        for (uint32_t i = 0; i < nk; ++i) ++cnt[key_kind[kb + i]];
        uint32_t at[3] = {0, cnt[0], cnt[0] + cnt[1]};
        for (uint32_t i = 0; i < nk; ++i) {
            const uint32_t kd = key_kind[kb + i];
            q->h_slot[kb + at[kd]++] = static_cast<uint16_t>(i | (kd << 14));
        }
    }
    q->device = ctx->device;
    q->n_keys = n_keys;
    q->prog_len = prog_len;
    q->kind_mask = kind_mask;
    q->n_units = corpus->n_units;
    q->row_words32 = 2 * ((n_keys + 63) / 64);
    const uint64_t matrix_words32 = std::max<uint64_t>(q->n_units * q->row_words32, 1);
    const uint64_t mask_words32 = std::max<uint64_t>(2 * ((q->n_units + 63) / 64), 1);
    const size_t cap_matrix_before = q->cap_matrix, cap_mask_before = q->cap_mask;
    if (!use_pinned) {
        CUDA_TRY(ensure_cap(q->d_keys, q->cap_keys, nbytes + kKeyPad));
        CUDA_TRY(ensure_cap(q->d_key_off, q->cap_key_off, (static_cast<uint64_t>(n_keys) + 1) * 8));
        CUDA_TRY(ensure_cap(q->d_kinds, q->cap_kinds, std::max<uint32_t>(n_keys, 1)));
        CUDA_TRY(ensure_cap(q->d_prog, q->cap_prog, std::max<uint32_t>(prog_len, 1) * sizeof(bsg_expr_op)));
        CUDA_TRY(ensure_cap(q->d_slot, q->cap_slot, std::max<uint32_t>(n_keys, 1) * sizeof(uint16_t)));
    }
    CUDA_TRY(ensure_cap(q->d_hashes, q->cap_hashes, std::max<uint64_t>(n_keys, 1) * 32));
    CUDA_TRY(ensure_cap(q->d_matrix32, q->cap_matrix, matrix_words32 * 4));
    CUDA_TRY(ensure_cap(q->d_mask32, q->cap_mask, mask_words32 * 4));
    // The kernels overwrite every word that carries a key / unit; pad words are never written, so
    // they only need zeroing when the buffers were reallocated or the output shape changed.
    const uint32_t groups = (n_keys + 31) / 32;  // 32-key words a row really carries (the rest is pad)
    const bool reshaped = q->zeroed_units != q->n_units || q->zeroed_row_words32 != q->row_words32 ||
                          q->zeroed_groups != groups || q->cap_matrix != cap_matrix_before ||
                          q->cap_mask != cap_mask_before;
    if (reshaped) {
        CUDA_TRY(cudaMemsetAsync(q->d_matrix32, 0, matrix_words32 * 4, s));
        CUDA_TRY(cudaMemsetAsync(q->d_mask32, 0, mask_words32 * 4, s));
        q->zeroed_units = q->n_units;
        q->zeroed_row_words32 = q->row_words32;
        q->zeroed_groups = groups;
    }
    if (use_pinned) {
        // one pinned block [keys + pad][offsets][kinds][program], mirrored on the device: ONE async copy
        const size_t o_off = (nbytes + kKeyPad + 15) & ~size_t(15);
        const size_t o_kind = o_off + (static_cast<size_t>(n_keys) + 1) * 8;
        const size_t o_prog = (o_kind + n_keys + 15) & ~size_t(15);
        const size_t o_slot = (o_prog + static_cast<size_t>(prog_len) * sizeof(bsg_expr_op) + 15) & ~size_t(15);
        const size_t used = o_slot + static_cast<size_t>(n_keys) * sizeof(uint16_t);
        const size_t need = used + 16;
        if (need > q->cap_pin) {
            if (q->h_pin) cudaFreeHost(q->h_pin);
            q->h_pin = nullptr;
            q->cap_pin = 0;
            CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&q->h_pin), need * 2, cudaHostAllocDefault));
            q->cap_pin = need * 2;
        }
        CUDA_TRY(ensure_cap(q->d_in, q->cap_in, need));
        if (nbytes) memcpy(q->h_pin, keys, nbytes);
        memset(q->h_pin + nbytes, 0, o_off - nbytes);
        if (n_keys) {
            memcpy(q->h_pin + o_off, key_off, (static_cast<size_t>(n_keys) + 1) * 8);
            memcpy(q->h_pin + o_kind, key_kind, n_keys);
        }
        if (prog_len) memcpy(q->h_pin + o_prog, prog, prog_len * sizeof(bsg_expr_op));
        if (n_keys) memcpy(q->h_pin + o_slot, q->h_slot.data(), n_keys * sizeof(uint16_t));
        CUDA_TRY(cudaMemcpyAsync(q->d_in, q->h_pin, used, cudaMemcpyHostToDevice, s));
        q->k_keys = q->d_in;
        q->k_key_off = reinterpret_cast<const uint64_t*>(q->d_in + o_off);
        q->k_kinds = q->d_in + o_kind;
        q->k_prog = reinterpret_cast<const bsg_expr_op*>(q->d_in + o_prog);
        q->k_slot = reinterpret_cast<const uint16_t*>(q->d_in + o_slot);
    } else {
        CUDA_TRY(cudaMemsetAsync(q->d_keys + nbytes, 0, kKeyPad, s));
        if (nbytes) CUDA_TRY(cudaMemcpyAsync(q->d_keys, keys, nbytes, cudaMemcpyHostToDevice, s));
        if (n_keys) {
            CUDA_TRY(cudaMemcpyAsync(q->d_key_off, key_off, (static_cast<uint64_t>(n_keys) + 1) * 8, cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(q->d_kinds, key_kind, n_keys, cudaMemcpyHostToDevice, s));
        }
        if (prog_len) CUDA_TRY(cudaMemcpyAsync(q->d_prog, prog, prog_len * sizeof(bsg_expr_op), cudaMemcpyHostToDevice, s));
        if (n_keys) CUDA_TRY(cudaMemcpyAsync(q->d_slot, q->h_slot.data(), n_keys * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
        q->k_slot = q->d_slot;
        q->k_keys = q->d_keys;
        q->k_key_off = q->d_key_off;
        q->k_kinds = q->d_kinds;
        q->k_prog = q->d_prog;
    }
    q->hashed = false;
    if (n_keys && !(use_pinned && ctx->fuse_hash)) {
        CUDA_TRY(launch_hash_keys(q->k_keys, q->k_key_off, n_keys, q->d_hashes, s));
        q->hashed = true;
    }
    q->k_matrix = q->d_matrix32;
    q->direct_out = false;
    return BSG_OK;
}

// bsg_probe() without a mask: let the probe kernels write the matrix straight into pinned host memory.
static int query_use_host_matrix(bsg_query* q) {
    const size_t bytes = std::max<uint64_t>(q->n_units * q->row_words32, 1) * 4;
    const uint32_t groups = (q->n_keys + 31) / 32;
    if (bytes > q->cap_out) {
        if (q->h_out) cudaFreeHost(q->h_out);
        q->h_out = nullptr;
        q->cap_out = 0;
        CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&q->h_out), bytes * 2, cudaHostAllocMapped));
        CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&q->h_out_dev), q->h_out, 0));
        q->cap_out = bytes * 2;
        q->out_units = ~0ull;
    }
    if (q->out_units != q->n_units || q->out_row_words32 != q->row_words32 || q->out_groups != groups) {
        memset(q->h_out, 0, bytes);  // pad words are never written by the kernels
        q->out_units = q->n_units;
        q->out_row_words32 = q->row_words32;
        q->out_groups = groups;
    }
    q->k_matrix = q->h_out_dev;
    return BSG_OK;
}

static int query_create_on(bsg_ctx* ctx, const bsg_corpus* corpus, const uint8_t* keys, const uint64_t* key_off,
                           uint32_t n_keys, const uint8_t* key_kind, const bsg_expr_op* prog, uint32_t prog_len,
                           cudaStream_t s, bsg_query** out) {
    if (!out) return fail(BSG_ERR_INVALID, "NULL argument");
    *out = nullptr;
    bsg_query* q = new (std::nothrow) bsg_query();
    if (!q) return fail(BSG_ERR_NOMEM, "query alloc");
    int rc = query_prepare_on(ctx, corpus, keys, key_off, n_keys, key_kind, prog, prog_len, s, q);
    if (rc) { bsg_query_free(q); return rc; }
    *out = q;
    return BSG_OK;
}

extern "C" int bsg_query_create(bsg_ctx* ctx, const bsg_corpus* corpus, const uint8_t* keys, const uint64_t* key_off,
                                uint32_t n_keys, const uint8_t* key_kind, const bsg_expr_op* prog, uint32_t prog_len,
                                bsg_query** out) {
    if (!ctx) return fail(BSG_ERR_INVALID, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = query_create_on(ctx, corpus, keys, key_off, n_keys, key_kind, prog, prog_len, ctx->cur_stream, out);
    if (rc) return rc;
    cudaError_t e = cudaStreamSynchronize(ctx->cur_stream);  // host buffers may be reused by the caller
    if (e != cudaSuccess) { bsg_query_free(*out); *out = nullptr; return fail(BSG_ERR_CUDA, "%s", cudaGetErrorString(e)); }
    return BSG_OK;
}

// Which staged kernel serves this corpus (measured on B200, profiles/r02_*): the tile ring wins on small units
// (several units per stage: the flush-shaped 1 000-row blocks) and is the only staged path for units of
// 75-190 KB (one stage per kind); the two-phase probe_staged2 kernel wins on units of 20-75 KB (merged
// 10 000-row blocks), where one unit fills a stage.
static int staged_variant_for(const bsg_ctx* ctx, const bsg_corpus* c) {
    if (ctx->probe_variant != 7) return ctx->probe_variant;
    if (c->n_staged == 0 || c->t_gather < c->n_gather) return 6;
    uint64_t bytes = 0;
    for (int k = 0; k < 3; ++k) bytes += c->staged_kind_bytes[k];
    return bytes / c->n_staged < 20 * 1024 ? 6 : 3;
}

// BSG_PROBE_AUTO: stream the staged units' bitsets through shared memory, or gather single words?
// staged traffic = every byte of the touched kinds; gather traffic ~ one 32 B sector per tested location
// (~3 on average for an absent key, k for a present one) + the descriptor.
static bool auto_uses_staged(const bsg_corpus* c, int variant, uint32_t kind_mask, uint32_t n_keys) {
    const uint32_t n_st = variant == 6 ? c->t_staged : c->n_staged;
    if (n_st == 0) return false;
    uint64_t staged_bytes = 0;
    for (int k = 0; k < 3; ++k)
        if (kind_mask & (1u << k)) staged_bytes += variant == 6 ? c->t_staged_kind_bytes[k] : c->staged_kind_bytes[k];
    return staged_bytes <= static_cast<uint64_t>(n_st) * n_keys * (4 * 32 + 32);
}

static int query_run_on(bsg_ctx* ctx, const bsg_corpus* c, bsg_query* q, int path, int want_matrix, cudaStream_t s,
                        const uint32_t* d_parent_mask32 = nullptr) {
    // d_parent_mask32 != nullptr: hierarchical stage — unit u is probed only if bit c->d_parent[u] of
    // the mask is set; its result is forced to "disqualified" otherwise.
    const uint32_t* d_parent = d_parent_mask32 ? c->d_parent : nullptr;
    if (d_parent_mask32 && !d_parent) return fail(BSG_ERR_INVALID, "corpus has no parents (bsg_corpus_set_parents)");
    if (!ctx || !c || !q) return fail(BSG_ERR_INVALID, "NULL argument");
    const bool matrix_only = (path & BSG_RUN_MATRIX_ONLY) != 0;
    // the matrix is always materialised (it is the tree kernel's input); a caller that wants ONLY the mask of a small
    // query lets the gather kernel short-circuit like evaluateBloomExpression does (the rows are then optimistic)
    const bool sc = !want_matrix && !matrix_only && q->prog_len > 0 && q->n_keys <= 32 && ctx->short_circuit;
    const bsg_expr_op* sc_prog = sc ? q->k_prog : nullptr;
    const uint32_t sc_len = sc ? q->prog_len : 0;
    path &= 0xff;
    if (q->n_units != c->n_units) return fail(BSG_ERR_INVALID, "query was created for a corpus of %llu units",
                                              (unsigned long long)q->n_units);
    if (path != BSG_PROBE_AUTO && path != BSG_PROBE_STAGED && path != BSG_PROBE_GATHER)
        return fail(BSG_ERR_INVALID, "unknown path %d", path);
    int launches = 0;
    const int variant = (ctx && c) ? staged_variant_for(ctx, c) : 3;
    if (q->n_keys && c->n_units && variant == 6) {
        // ---- tile ring (probe_tiles_kernel) for the units a tile can hold, gather kernel for the rest ----
        bool use_staged = c->t_staged > 0;
        if (path == BSG_PROBE_GATHER) use_staged = false;
        if (path == BSG_PROBE_AUTO && use_staged) use_staged = auto_uses_staged(c, 6, q->kind_mask, q->n_keys);
        // deferred hashing (bsg_probe path): every CTA of the tile kernel hashes the batch into its own shared
        // memory when that kernel is the only consumer of the hashes, else a hash_keys_kernel launch now
        const bool fuse = !q->hashed && use_staged && c->t_gather == 0;
        if (!q->hashed && !fuse) {
            CUDA_TRY(launch_hash_keys(q->k_keys, q->k_key_off, q->n_keys, q->d_hashes, s));
            q->hashed = true;
            ++launches;
        }
        if (use_staged) {
            ProbeTilesPlan plan;
            plan.shape = ctx->tiles_shape;
            plan.pdl = ctx->pdl;
            plan.parts = c->t_parts;
            plan.units_cap = c->t_units_cap;
            plan.stage_data_bytes = c->t_data_cap;
            plan.fuse_keys = fuse ? q->k_keys : nullptr;
            plan.fuse_key_off = fuse ? q->k_key_off : nullptr;
            const uint32_t nk = std::min<uint32_t>(q->n_keys, kProbeMaxKeysPerPass);
            const uint64_t stage_bytes = tile_header_bytes(plan.units_cap) + plan.stage_data_bytes;
            int max_stages = kProbeMaxStages;
            if (ctx->max_stages > 0 && ctx->max_stages < max_stages) max_stages = ctx->max_stages;
            const uint32_t fixed = tiles_fixed_smem(c->t_units_cap, nk);
            plan.n_stages = static_cast<int>(std::min<uint64_t>(max_stages, (tiles_cta_smem(ctx) - fixed) / stage_bytes));
            if (plan.n_stages < 1) return fail(BSG_ERR_INVALID, "internal: tile does not fit shared memory");
            plan.smem_bytes = fixed + plan.n_stages * stage_bytes;
            plan.grid = static_cast<int>(std::min<uint64_t>(
                c->t_items, static_cast<uint64_t>(ctx->sm_count) * (1024 / probe_tiles_threads(plan.shape))));
            const TileRec* tiles = c->d_tiles;
            const uint32_t* d_n_items = nullptr;
            if (d_parent) {  // keep the tiles with a unit whose parent survived (compacted on the device)
                CUDA_TRY(ensure_cap(q->d_tiles_c, q->cap_tiles_c,
                                    static_cast<size_t>(c->t_items) * c->t_parts * sizeof(TileRec)));
                if (!q->d_n_rows) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&q->d_n_rows), 4));
                CUDA_TRY(launch_compact_tiles(c->d_tiles, c->t_items, c->t_parts, d_parent, d_parent_mask32, q->d_tiles_c,
                                              q->d_n_rows, s));
                ++launches;
                tiles = q->d_tiles_c;
                d_n_items = q->d_n_rows;
            }
            for (uint32_t kb = 0; kb < q->n_keys; kb += kProbeMaxKeysPerPass) {
                const uint32_t nk = std::min<uint32_t>(kProbeMaxKeysPerPass, q->n_keys - kb);
                CUDA_TRY(launch_probe_tiles(plan, tiles, c->t_items, d_n_items, c->d_words, q->d_hashes, q->k_slot, kb, nk,
                                            q->kind_mask, q->k_matrix, q->row_words32, s, ctx->d_trace, ctx->trace_slots));
                ++launches;
            }
        } else if (c->t_staged) {
            CUDA_TRY(launch_probe_gather(c->d_udesc, c->d_words, c->d_t_staged_list, c->t_staged, q->d_hashes,
                                         q->k_kinds, q->n_keys, q->k_matrix, q->row_words32, s, d_parent,
                                         d_parent_mask32, sc_prog, sc_len));
            ++launches;
        }
        if (c->t_gather) {
            CUDA_TRY(launch_probe_gather(c->d_udesc, c->d_words, c->d_t_gather_list, c->t_gather, q->d_hashes,
                                         q->k_kinds, q->n_keys, q->k_matrix, q->row_words32, s, d_parent,
                                         d_parent_mask32, sc_prog, sc_len));
            ++launches;
        }
    } else if (q->n_keys && c->n_units) {
        // --- choose the data path for the stageable units ---
        bool use_staged = c->n_staged > 0;
        if (path == BSG_PROBE_GATHER) use_staged = false;
        if (path == BSG_PROBE_AUTO && use_staged) use_staged = auto_uses_staged(c, variant, q->kind_mask, q->n_keys);
        // deferred hashing (bsg_probe path): fused into the two-phase staged kernel when that kernel is
        // the only consumer of the hashes, else a hash_keys_kernel launch now
        const bool fuse = !q->hashed && use_staged && c->n_gather == 0 && variant != 0;
        if (!q->hashed && !fuse) {
            CUDA_TRY(launch_hash_keys(q->k_keys, q->k_key_off, q->n_keys, q->d_hashes, s));
            q->hashed = true;
            ++launches;
        }
        if (use_staged) {
            ProbeStagedPlan plan;
            plan.variant = variant;
            plan.relax_sleep_ns = static_cast<uint32_t>(ctx->relax_sleep_ns);
            plan.pdl = ctx->pdl;
            plan.fuse_keys = nullptr;
            plan.fuse_key_off = nullptr;
            plan.fuse_scratch = nullptr;
            const uint64_t prefix = plan.variant ? kProbe2SmemPrefixBytes : kProbeSmemPrefixBytes;
            const uint64_t budget = static_cast<uint64_t>(ctx->max_smem_optin) - prefix;
            plan.stage_data_bytes = std::max<uint32_t>(c->stage_cap_bytes, 16);
            const uint64_t stage_bytes =
                (plan.variant ? kProbeStage2HeaderBytes : kProbeStageHeaderBytes) + plan.stage_data_bytes;
            int max_stages = kProbeMaxStages;
            if (ctx->max_stages > 0 && ctx->max_stages < max_stages) max_stages = ctx->max_stages;
            plan.n_stages = static_cast<int>(std::min<uint64_t>(max_stages, budget / stage_bytes));
            if (plan.n_stages < 1) return fail(BSG_ERR_INVALID, "internal: stage does not fit shared memory");
            plan.smem_bytes = prefix + plan.n_stages * stage_bytes;
            plan.grid = static_cast<int>(std::min<uint64_t>(c->n_staged, ctx->sm_count));
            plan.warps = ctx->probe_warps;
            // time for the whole chip to stream one stage per SM at ~6.5 TB/s, capped at 2 us
            {
                const double ns = static_cast<double>(stage_bytes) * plan.grid / 6500.0;
                plan.stagger_ns = ctx->stagger_pct < 0 ? 0u
                                                       : static_cast<uint32_t>(std::min(2000.0, ns * ctx->stagger_pct / 100.0));
            }
            if (fuse) {
                CUDA_TRY(ensure_cap(q->d_hash_scratch, q->cap_hash_scratch,
                                    static_cast<size_t>(plan.grid) * kProbeMaxKeysPerPass * 32));
                plan.fuse_keys = q->k_keys;
                plan.fuse_key_off = q->k_key_off;
                plan.fuse_scratch = q->d_hash_scratch;
            }
            const StageRow* rows = c->d_stab;
            const uint32_t* d_n_rows = nullptr;
            if (d_parent) {  // compact the surviving units' stage rows on the device
                CUDA_TRY(ensure_cap(q->d_rows, q->cap_rows, static_cast<size_t>(c->n_staged) * sizeof(StageRow)));
                if (!q->d_n_rows) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&q->d_n_rows), 4));
                CUDA_TRY(launch_compact_rows(c->d_stab, c->n_staged, d_parent, d_parent_mask32, q->d_rows, q->d_n_rows, s));
                ++launches;
                rows = q->d_rows;
                d_n_rows = q->d_n_rows;
            }
            for (uint32_t kb = 0; kb < q->n_keys; kb += kProbeMaxKeysPerPass) {
                const uint32_t nk = std::min<uint32_t>(kProbeMaxKeysPerPass, q->n_keys - kb);
                CUDA_TRY(launch_probe_staged(plan, rows, c->n_staged, c->d_words, q->d_hashes, q->k_kinds, kb, nk,
                                             q->kind_mask, q->k_matrix, q->row_words32, s, ctx->d_trace,
                                             ctx->trace_slots, d_n_rows));
                ++launches;
            }
        } else if (c->n_staged) {
            CUDA_TRY(launch_probe_gather(c->d_udesc, c->d_words, c->d_staged_list, c->n_staged, q->d_hashes,
                                         q->k_kinds, q->n_keys, q->k_matrix, q->row_words32, s, d_parent,
                                         d_parent_mask32, sc_prog, sc_len));
            ++launches;
        }
        if (c->n_gather) {
            CUDA_TRY(launch_probe_gather(c->d_udesc, c->d_words, c->d_gather_list, c->n_gather, q->d_hashes,
                                         q->k_kinds, q->n_keys, q->k_matrix, q->row_words32, s, d_parent,
                                         d_parent_mask32, sc_prog, sc_len));
            ++launches;
        }
    }
    if (c->n_units && !matrix_only) {
        if (q->prog_len) {
            CUDA_TRY(launch_tree_eval(q->k_matrix, q->row_words32, c->n_units, q->k_prog, q->prog_len, q->d_mask32, s,
                                      d_parent, d_parent_mask32));
        } else if (d_parent) {
            CUDA_TRY(launch_parent_mask(q->d_mask32, c->n_units, d_parent, d_parent_mask32, s));
        } else {
            CUDA_TRY(launch_fill_mask(q->d_mask32, c->n_units, s));
        }
        ++launches;
        if (c->d_bad32) {  // a unit whose section failed to parse is an error, not a candidate
            CUDA_TRY(launch_mask_andnot(q->d_mask32, c->d_bad32, c->n_units, s));
            ++launches;
        }
    }
    q->last_launches = launches;
    return BSG_OK;
}

extern "C" int bsg_query_run(bsg_ctx* ctx, const bsg_corpus* corpus, bsg_query* q, int path, int want_matrix) {
    if (!ctx) return fail(BSG_ERR_INVALID, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return query_run_on(ctx, corpus, q, path, want_matrix, ctx->cur_stream);
}

// D2H through the query's pinned block (true async DMA), then memcpy into the caller's buffers.
static int query_fetch_pinned(bsg_ctx* ctx, bsg_query* q, uint64_t n_units, uint64_t* out_matrix, uint64_t* out_mask,
                              cudaStream_t s, std::chrono::steady_clock::time_point* t_synced = nullptr) {
    const size_t mbytes = (out_matrix && q->n_keys) ? n_units * q->row_words32 * 4 : 0;
    const size_t kbytes = out_mask ? ((n_units + 63) / 64) * 8 : 0;
    if (q->k_matrix != q->d_matrix32) {  // zero copy: the rows are already in pinned host memory
        CUDA_TRY(stream_wait(ctx, s));
        if (t_synced) *t_synced = std::chrono::steady_clock::now();
        if (mbytes && !q->direct_out) memcpy(out_matrix, q->h_out, mbytes);
        return BSG_OK;
    }
    const size_t need = mbytes + kbytes + 16;
    if (need > q->cap_pin) {  // grow, keeping nothing (inputs were already consumed by the copies on s)
        CUDA_TRY(cudaStreamSynchronize(s));
        if (q->h_pin) cudaFreeHost(q->h_pin);
        q->h_pin = nullptr;
        q->cap_pin = 0;
        CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&q->h_pin), need * 2, cudaHostAllocDefault));
        q->cap_pin = need * 2;
    } else {
        // the input staging area is reused for the outputs: the H2D copies must have been consumed
        // -> they precede the kernels on s, and the D2H copies below follow the kernels on s.
    }
    if (mbytes) CUDA_TRY(cudaMemcpyAsync(q->h_pin, q->d_matrix32, mbytes, cudaMemcpyDeviceToHost, s));
    if (kbytes) CUDA_TRY(cudaMemcpyAsync(q->h_pin + mbytes, q->d_mask32, kbytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(stream_wait(ctx, s));
    if (t_synced) *t_synced = std::chrono::steady_clock::now();
    if (mbytes) memcpy(out_matrix, q->h_pin, mbytes);
    if (kbytes) memcpy(out_mask, q->h_pin + mbytes, kbytes);
    return BSG_OK;
}

static int query_fetch_on(bsg_query* q, uint64_t n_units, uint64_t* out_matrix, uint64_t* out_mask, cudaStream_t s) {
    if (n_units != q->n_units) return fail(BSG_ERR_INVALID, "n_units mismatch");
    if (out_matrix && q->n_keys && n_units)
        CUDA_TRY(cudaMemcpyAsync(out_matrix, q->d_matrix32, n_units * q->row_words32 * 4, cudaMemcpyDeviceToHost, s));
    if (out_mask && n_units)
        CUDA_TRY(cudaMemcpyAsync(out_mask, q->d_mask32, ((n_units + 63) / 64) * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return BSG_OK;
}

extern "C" int bsg_query_fetch(bsg_ctx* ctx, bsg_query* q, uint64_t n_units, uint64_t* out_matrix, uint64_t* out_mask) {
    if (!ctx || !q) return fail(BSG_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return query_fetch_on(q, n_units, out_matrix, out_mask, ctx->cur_stream);
}

extern "C" int bsg_probe(bsg_ctx* ctx, const bsg_corpus* corpus, const uint8_t* keys, const uint64_t* key_off,
                         uint32_t n_keys, const uint8_t* key_kind, const bsg_expr_op* prog, uint32_t prog_len,
                         uint64_t* out_matrix, uint64_t* out_mask) {
    if (!ctx || !corpus) return fail(BSG_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = pool_get(ctx);
    if (!s) return fail(BSG_ERR_CUDA, "stream create failed");
    bsg_query* q = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (!ctx->scratch_pool.empty()) { q = ctx->scratch_pool.back(); ctx->scratch_pool.pop_back(); }
    }
    if (!q) q = new (std::nothrow) bsg_query();
    int rc = q ? BSG_OK : fail(BSG_ERR_NOMEM, "query alloc");
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    if (rc == BSG_OK) rc = query_prepare_on(ctx, corpus, keys, key_off, n_keys, key_kind, prog, prog_len, s, q, true);
    // zero copy pays while the rows trickle out under the kernel's own runtime; a matrix of many MB would
    // instead hold the SMs at PCIe speed, so large outputs keep the copy-engine path (D2H after the kernel)
    constexpr uint64_t kZeroCopyMaxBytes = 8ull << 20;
    if (rc == BSG_OK && ctx->zero_copy && out_matrix && !out_mask && n_keys && corpus->n_units &&
        corpus->n_units * q->row_words32 * 4ull <= kZeroCopyMaxBytes) {
        // a caller buffer from bsg_host_alloc takes the rows directly (the staged kernels write every word of a
        // row, pad included; units on the gather kernel do not, so those corpora keep the staging block)
        uint32_t* direct = nullptr;
        const int v = staged_variant_for(ctx, corpus);
        const bool all_staged = v == 6 ? (corpus->t_gather == 0 && corpus->t_staged > 0) : (corpus->n_gather == 0 && corpus->n_staged > 0);
        if (all_staged && v != 0 && auto_uses_staged(corpus, v, q->kind_mask, n_keys))
            direct = host_buf_device_ptr(ctx, out_matrix, corpus->n_units * q->row_words32 * 4ull);
        if (direct) { q->k_matrix = direct; q->direct_out = true; }
        else rc = query_use_host_matrix(q);
    }
    const auto t1 = clk::now();
    // the mask kernel is skipped when the caller wants no mask
    const int path = BSG_PROBE_AUTO | (out_mask ? 0 : BSG_RUN_MATRIX_ONLY);
    if (rc == BSG_OK) rc = query_run_on(ctx, corpus, q, path, out_matrix != nullptr, s);
    const auto t2 = clk::now();
    clk::time_point t3 = t2;
    if (rc == BSG_OK) rc = query_fetch_pinned(ctx, q, corpus->n_units, out_matrix, out_mask, s, ctx->timing ? &t3 : nullptr);
    else cudaStreamSynchronize(s);
    if (ctx->timing) {
        const auto t4 = clk::now();
        auto ns = [](clk::time_point a, clk::time_point b) {
            return static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count());
        };
        ctx->t_calls += 1; ctx->t_prepare += ns(t0, t1); ctx->t_run += ns(t1, t2); ctx->t_wait += ns(t2, t3);
        ctx->t_copyout += ns(t3, t4);
    }
    if (q) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->scratch_pool.push_back(q);
    }
    pool_put(ctx, s);
    return rc;
}

// Several queries in ONE pass over the corpus (concurrent queries merged into one launch).
// The staged probe streams every filter byte of the corpus once per pass of up to 1 024 keys whatever the number
// of keys, so the keys of many small queries (the reference's typical query has a handful of leaves) ride along
// for free; each query keeps its own expression and gets its own candidate mask.
extern "C" int bsg_probe_multi(bsg_ctx* ctx, const bsg_corpus* corpus, const uint8_t* keys, const uint64_t* key_off,
                               uint32_t n_keys, const uint8_t* key_kind, uint32_t n_queries,
                               const uint32_t* query_key_begin, const bsg_expr_op* progs, const uint32_t* prog_begin,
                               uint64_t* out_masks) {
    if (!ctx || !corpus || !query_key_begin || !prog_begin || !out_masks || n_queries == 0)
        return fail(BSG_ERR_INVALID, "NULL argument");
    if (n_queries > 65535u) return fail(BSG_ERR_INVALID, "at most 65535 queries per call");
    if (query_key_begin[0] != 0 || query_key_begin[n_queries] != n_keys || prog_begin[0] != 0)
        return fail(BSG_ERR_INVALID, "query_key_begin / prog_begin must be CSR arrays covering all keys / ops");
    const uint32_t total_ops = prog_begin[n_queries];
    if (total_ops && !progs) return fail(BSG_ERR_INVALID, "progs is NULL");
    // leaves are query-local in the caller's programs and batch-global on the device; the CSR of the programs
    // rides behind the ops in the same upload (two uint32 per pseudo-op)
    const uint32_t tail_ops = (n_queries + 2u) / 2u;
    std::vector<bsg_expr_op> gprog(static_cast<size_t>(total_ops) + tail_ops, bsg_expr_op{0u, 0u});
    for (uint32_t j = 0; j < n_queries; ++j) {
        const uint32_t kb = query_key_begin[j], ke = query_key_begin[j + 1], pb = prog_begin[j], pe = prog_begin[j + 1];
        if (ke < kb || ke > n_keys || pe < pb || pe > total_ops)
            return fail(BSG_ERR_INVALID, "query %u: key / op range not monotone", j);
        if (pe > pb) {
            int rc = validate_program(progs + pb, pe - pb, ke - kb);
            if (rc) return rc;
        }
        for (uint32_t pc = pb; pc < pe; ++pc) {
            gprog[pc] = progs[pc];
            if (progs[pc].op == BSG_OP_LEAF) gprog[pc].arg += kb;
        }
    }
    memcpy(gprog.data() + total_ops, prog_begin, (static_cast<size_t>(n_queries) + 1) * sizeof(uint32_t));
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = pool_get(ctx);
    if (!s) return fail(BSG_ERR_CUDA, "stream create failed");
    bsg_query* q = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (!ctx->scratch_pool.empty()) { q = ctx->scratch_pool.back(); ctx->scratch_pool.pop_back(); }
    }
    if (!q) q = new (std::nothrow) bsg_query();
    int rc = q ? BSG_OK : fail(BSG_ERR_NOMEM, "query alloc");
    if (rc == BSG_OK)
        rc = query_prepare_on(ctx, corpus, keys, key_off, n_keys, key_kind, gprog.data(),
                              static_cast<uint32_t>(gprog.size()), s, q, true, true);
    const uint64_t n_units = corpus->n_units;
    const uint64_t mask_words32 = 2 * ((n_units + 63) / 64);
    const size_t out_bytes = static_cast<size_t>(n_queries) * mask_words32 * 4;
    auto run = [&]() -> int {
        int r = query_run_on(ctx, corpus, q, BSG_PROBE_AUTO | BSG_RUN_MATRIX_ONLY, 1, s);
        if (r) return r;
        if (n_units == 0) return BSG_OK;
        CUDA_TRY(ensure_cap(q->d_multi_mask32, q->cap_multi_mask, std::max<size_t>(out_bytes, 16)));
        CUDA_TRY(launch_tree_eval_multi(q->k_matrix, q->row_words32, n_units, q->k_prog,
                                        reinterpret_cast<const uint32_t*>(q->k_prog + total_ops), n_queries,
                                        q->d_multi_mask32, mask_words32, corpus->d_bad32, s));
        q->last_launches += 1;
        if (out_bytes + 16 > q->cap_pin) {  // the inputs in the pinned block must have been consumed first
            CUDA_TRY(cudaStreamSynchronize(s));
            if (q->h_pin) cudaFreeHost(q->h_pin);
            q->h_pin = nullptr;
            q->cap_pin = 0;
            CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&q->h_pin), (out_bytes + 16) * 2, cudaHostAllocDefault));
            q->cap_pin = (out_bytes + 16) * 2;
        }
        CUDA_TRY(cudaMemcpyAsync(q->h_pin, q->d_multi_mask32, out_bytes, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(stream_wait(ctx, s));
        memcpy(out_masks, q->h_pin, out_bytes);
        return BSG_OK;
    };
    if (rc == BSG_OK) rc = run();
    if (rc != BSG_OK) cudaStreamSynchronize(s);
    if (q) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->scratch_pool.push_back(q);
    }
    pool_put(ctx, s);
    return rc;
}

extern "C" int bsg_corpus_set_parents(bsg_ctx* ctx, bsg_corpus* corpus, const uint32_t* parent, uint64_t n_units,
                                      uint64_t n_parent_units) {
    if (!ctx || !corpus || (n_units && !parent)) return fail(BSG_ERR_INVALID, "NULL argument");
    if (n_units != corpus->n_units) return fail(BSG_ERR_INVALID, "parent array must have one entry per unit");
    for (uint64_t u = 0; u < n_units; ++u)
        if (parent[u] >= n_parent_units) return fail(BSG_ERR_INVALID, "unit %llu: parent %u out of range", (unsigned long long)u, parent[u]);
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaFree(corpus->d_parent);
    corpus->d_parent = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&corpus->d_parent), std::max<uint64_t>(n_units, 1) * 4));
    if (n_units) CUDA_TRY(cudaMemcpy(corpus->d_parent, parent, n_units * 4, cudaMemcpyHostToDevice));
    corpus->n_parents = n_parent_units;
    return BSG_OK;
}

static bsg_query* scratch_get(bsg_ctx* ctx) {
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (!ctx->scratch_pool.empty()) {
            bsg_query* q = ctx->scratch_pool.back();
            ctx->scratch_pool.pop_back();
            return q;
        }
    }
    return new (std::nothrow) bsg_query();
}
static void scratch_put(bsg_ctx* ctx, bsg_query* q) {
    if (!q) return;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->scratch_pool.push_back(q);
}

extern "C" int bsg_probe_hierarchical(bsg_ctx* ctx, const bsg_corpus* files, const bsg_corpus* blocks,
                                      const uint8_t* keys, const uint64_t* key_off, uint32_t n_keys,
                                      const uint8_t* key_kind, const bsg_expr_op* prog, uint32_t prog_len,
                                      uint64_t* out_file_mask, uint64_t* out_block_mask) {
    if (!ctx || !files || !blocks) return fail(BSG_ERR_INVALID, "NULL argument");
    if (!blocks->d_parent || blocks->n_parents != files->n_units)
        return fail(BSG_ERR_INVALID, "blocks corpus has no parents into this files corpus (bsg_corpus_set_parents)");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = pool_get(ctx);
    if (!s) return fail(BSG_ERR_CUDA, "stream create failed");
    bsg_query* qf = scratch_get(ctx);
    bsg_query* qb = scratch_get(ctx);
    int rc = (qf && qb) ? BSG_OK : fail(BSG_ERR_NOMEM, "query alloc");
    // stage 1: file level (query_exec.go:399-406)
    if (rc == BSG_OK) rc = query_prepare_on(ctx, files, keys, key_off, n_keys, key_kind, prog, prog_len, s, qf, true);
    if (rc == BSG_OK) rc = query_run_on(ctx, files, qf, BSG_PROBE_AUTO, 0, s);
    // stage 2: block level, only blocks of surviving files (query_exec.go:572-615)
    if (rc == BSG_OK) rc = query_prepare_on(ctx, blocks, keys, key_off, n_keys, key_kind, prog, prog_len, s, qb, true);
    if (rc == BSG_OK) rc = query_run_on(ctx, blocks, qb, BSG_PROBE_AUTO, 0, s, qf->d_mask32);
    if (rc == BSG_OK && out_file_mask)
        rc = query_fetch_on(qf, files->n_units, nullptr, out_file_mask, s);
    if (rc == BSG_OK) rc = query_fetch_pinned(ctx, qb, blocks->n_units, nullptr, out_block_mask, s);
    else cudaStreamSynchronize(s);
    scratch_put(ctx, qf);
    scratch_put(ctx, qb);
    pool_put(ctx, s);
    return rc;
}

extern "C" int bsg_query_run_child(bsg_ctx* ctx, const bsg_corpus* blocks, bsg_query* q, const bsg_query* parent_q, int path) {
    if (!ctx || !blocks || !q || !parent_q) return fail(BSG_ERR_INVALID, "NULL argument");
    if (!blocks->d_parent || blocks->n_parents != parent_q->n_units)
        return fail(BSG_ERR_INVALID, "blocks corpus has no parents into the parent query's corpus (bsg_corpus_set_parents)");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return query_run_on(ctx, blocks, q, path, 0, ctx->cur_stream, parent_q->d_mask32);
}

extern "C" const uint64_t* bsg_query_device_mask(const bsg_query* q) {
    return q ? reinterpret_cast<const uint64_t*>(q->d_mask32) : nullptr;
}

// Sharded hierarchical probe: bsg_probe_hierarchical on this rank's shard, then the per-rank block masks are
// all-gathered on the device (bsg_comm.cpp) and copied out once.
extern "C" int bsg_allgather_masks_device(bsg_ctx* ctx, const uint64_t* d_local, uint64_t n_words, uint64_t* d_all);
extern "C" int bsg_comm_info(bsg_ctx* ctx, int* rank, int* world, int* peer_memory, uint64_t* last_nvlink_bytes);
extern "C" int bsg_comm_alloc(bsg_ctx* ctx, size_t bytes, void** out_dev);
extern "C" int bsg_comm_free(bsg_ctx* ctx, void* dev);

extern "C" int bsg_probe_hierarchical_gather(bsg_ctx* ctx, const bsg_corpus* files, const bsg_corpus* blocks,
                                             const uint8_t* keys, const uint64_t* key_off, uint32_t n_keys,
                                             const uint8_t* key_kind, const bsg_expr_op* prog, uint32_t prog_len,
                                             uint64_t mask_words, uint64_t* out_all_block_masks) {
    if (!ctx || !files || !blocks || !out_all_block_masks) return fail(BSG_ERR_INVALID, "NULL argument");
    if (!blocks->d_parent || blocks->n_parents != files->n_units)
        return fail(BSG_ERR_INVALID, "blocks corpus has no parents into this files corpus (bsg_corpus_set_parents)");
    if (mask_words < (blocks->n_units + 63) / 64) return fail(BSG_ERR_INVALID, "mask_words smaller than this rank's mask");
    int world = 1;
    int rc = bsg_comm_info(ctx, nullptr, &world, nullptr, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    // gather landing zone (symmetric) + padded local mask, cached on the ctx; grows collectively
    const size_t need = (static_cast<size_t>(world) + 1) * mask_words * 8;
    if (need > ctx->gather_cap) {
        if (ctx->d_gather) { rc = bsg_comm_free(ctx, ctx->d_gather); if (rc) return rc; }
        ctx->d_gather = nullptr;
        ctx->gather_cap = 0;
        void* p = nullptr;
        rc = bsg_comm_alloc(ctx, need * 2, &p);
        if (rc) return rc;
        ctx->d_gather = static_cast<uint64_t*>(p);
        ctx->gather_cap = need * 2;
        if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
        ctx->h_gather = nullptr;
        CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_gather), need * 2, cudaHostAllocDefault));
    }
    uint64_t* d_all = ctx->d_gather;
    uint64_t* d_local = ctx->d_gather + static_cast<size_t>(world) * mask_words;
    cudaStream_t s = ctx->cur_stream;   // collectives are ordered on the ctx stream
    bsg_query* qf = scratch_get(ctx);
    bsg_query* qb = scratch_get(ctx);
    rc = (qf && qb) ? BSG_OK : fail(BSG_ERR_NOMEM, "query alloc");
    if (rc == BSG_OK) rc = query_prepare_on(ctx, files, keys, key_off, n_keys, key_kind, prog, prog_len, s, qf, true);
    if (rc == BSG_OK) rc = query_run_on(ctx, files, qf, BSG_PROBE_AUTO, 0, s);
    if (rc == BSG_OK) rc = query_prepare_on(ctx, blocks, keys, key_off, n_keys, key_kind, prog, prog_len, s, qb, true);
    if (rc == BSG_OK) rc = query_run_on(ctx, blocks, qb, BSG_PROBE_AUTO, 0, s, qf->d_mask32);
    if (rc == BSG_OK) {
        const size_t mine = ((blocks->n_units + 63) / 64) * 8;
        cudaError_t e = cudaMemsetAsync(d_local, 0, mask_words * 8, s);
        if (e == cudaSuccess && mine) e = cudaMemcpyAsync(d_local, qb->d_mask32, mine, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) rc = fail(BSG_ERR_CUDA, "%s", cudaGetErrorString(e));
    }
    if (rc == BSG_OK) rc = bsg_allgather_masks_device(ctx, d_local, mask_words, d_all);
    if (rc == BSG_OK) {
        cudaError_t e = cudaMemcpyAsync(ctx->h_gather, d_all, static_cast<size_t>(world) * mask_words * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = fail(BSG_ERR_CUDA, "%s", cudaGetErrorString(e));
        else memcpy(out_all_block_masks, ctx->h_gather, static_cast<size_t>(world) * mask_words * 8);
    } else {
        cudaStreamSynchronize(s);
    }
    scratch_put(ctx, qf);
    scratch_put(ctx, qb);
    return rc;
}

// ---- profiling-only hooks (not part of include/bloomgpu.h): per-CTA timeline of the staged probe
extern "C" int bsg_debug_trace_enable(bsg_ctx* ctx, uint32_t slots) {
    if (!ctx) return fail(BSG_ERR_INVALID, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaFree(ctx->d_trace);
    ctx->d_trace = nullptr;
    ctx->trace_slots = slots;
    if (slots) {
        const size_t bytes = static_cast<size_t>(ctx->sm_count) * slots * 8;
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ctx->d_trace), bytes));
        CUDA_TRY(cudaMemset(ctx->d_trace, 0, bytes));
    }
    return BSG_OK;
}
extern "C" int bsg_debug_trace_read(bsg_ctx* ctx, uint64_t* out /* sm_count * slots */) {
    if (!ctx || !ctx->d_trace || !out) return fail(BSG_ERR_INVALID, "trace not enabled");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(out, ctx->d_trace, static_cast<size_t>(ctx->sm_count) * ctx->trace_slots * 8,
                        cudaMemcpyDeviceToHost));
    return BSG_OK;
}

// Measurement helper: enqueue `steps` runs from C (no interpreter between launches), cycling
// over n (corpus, query) pairs so consecutive steps touch different HBM.  n_streams > 1 issues
// consecutive steps round-robin on that many internal streams, forked from / joined to the ctx
// stream with events, so independent batches overlap tail-to-head exactly as concurrent
// bsg_probe() callers (one pool stream each) do.
extern "C" int bsg_debug_run_cycle(bsg_ctx* ctx, bsg_corpus* const* corpora, bsg_query* const* queries, uint32_t n,
                                   uint32_t steps, int path, uint32_t n_streams) {
    if (!ctx || !corpora || !queries || n == 0) return fail(BSG_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (n_streams <= 1) {
        for (uint32_t i = 0; i < steps; ++i) {
            int rc = query_run_on(ctx, corpora[i % n], queries[i % n], path, 1, ctx->cur_stream);
            if (rc) return rc;
        }
        return BSG_OK;
    }
    if (n_streams > 8) n_streams = 8;
    if (n_streams > n) n_streams = n;  // a query object must never be in flight twice
    while (ctx->aux_streams.size() < n_streams) {
        cudaStream_t st = nullptr;
        CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ctx->aux_streams.push_back(st);
        cudaEvent_t ev = nullptr;
        CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->aux_events.push_back(ev);
    }
    if (!ctx->fork_event) CUDA_TRY(cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(ctx->fork_event, ctx->cur_stream));
    for (uint32_t k = 0; k < n_streams; ++k) CUDA_TRY(cudaStreamWaitEvent(ctx->aux_streams[k], ctx->fork_event, 0));
    for (uint32_t i = 0; i < steps; ++i) {
        // step i uses pair i % n on stream i % n_streams; with n a multiple of n_streams a pair is
        // always replayed on the same stream, so its own runs stay ordered
        int rc = query_run_on(ctx, corpora[i % n], queries[i % n], path, 1, ctx->aux_streams[i % n_streams]);
        if (rc) return rc;
    }
    for (uint32_t k = 0; k < n_streams; ++k) {
        CUDA_TRY(cudaEventRecord(ctx->aux_events[k], ctx->aux_streams[k]));
        CUDA_TRY(cudaStreamWaitEvent(ctx->cur_stream, ctx->aux_events[k], 0));
    }
    return BSG_OK;
}

// reporting helper: which staged kernel a big-batch probe of this corpus launches, and how the corpus was cut
extern "C" int bsg_debug_probe_kernel_name(bsg_ctx* ctx, const bsg_corpus* c, char* buf, size_t n) {
    if (!ctx || !c || !buf || n == 0) return fail(BSG_ERR_INVALID, "NULL argument");
    const int variant = staged_variant_for(ctx, c);
    if (variant == 6)
        snprintf(buf, n, "%s, %s mode, %u item(s) x %u tile(s), <= %u unit(s) and %u data bytes per tile; %u unit(s) on the gather kernel",
                 probe_tiles_shape_name(ctx->tiles_shape), c->t_parts == 1 ? "UNIT" : "KIND", c->t_items, c->t_parts,
                 c->t_units_cap, c->t_data_cap, c->t_gather);
    else if (variant == 0)
        snprintf(buf, n, "probe_staged_kernel (one phase)");
    else
        snprintf(buf, n, "probe_staged2_kernel shape %d (<16,2,3,16,4> by default; one team of 16 B warps when the ring has < 4 stages), %u staged unit(s), %u on the gather kernel", variant, c->n_staged, c->n_gather);
    return BSG_OK;
}

// test / bench helper: synchronise the ctx stream, then copy device memory (e.g. a bsg_comm_alloc buffer) to the host
extern "C" int bsg_debug_memcpy_d2h(bsg_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    if (!ctx || !dst_host || !src_dev) return fail(BSG_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->cur_stream));
    CUDA_TRY(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
    return BSG_OK;
}

extern "C" float bsg_debug_last_build_kernel_ms(bsg_ctx* ctx) { return ctx ? ctx->last_build_kernel_ms : 0.f; }

// ---- hooks for bsg_comm.cpp (keeps bsg_ctx's layout private to this file) ----
extern "C" int bsg_ctx_device_internal(bsg_ctx* ctx) { return ctx->device; }
extern "C" void** bsg_ctx_comm_slot_internal(bsg_ctx* ctx) { return &ctx->comm; }
extern "C" void* bsg_ctx_stream_internal(bsg_ctx* ctx) { return ctx->cur_stream; }
extern "C" int bsg_set_last_error_internal(int code, const char* msg) { return fail(code, "%s", msg); }
