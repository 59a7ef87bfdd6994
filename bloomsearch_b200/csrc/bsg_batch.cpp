// Query batcher ("batch concurrent queries into one launch").
//
// The reference runs every query on its own goroutines and re-reads the filters per query
// (query_exec.go:201-433); on the GPU the filters are resident and ONE pass over them can answer up to
// 1 024 keys, so concurrent queries against the same corpus are worth merging.  This is a group commit:
// the first caller to arrive while no launch is in flight becomes the leader of a batch and launches at
// once (an idle system adds no latency); callers that arrive while a launch is running join the next
// batch, whose leader launches as soon as the running one ends.  The
// leader concatenates the members' keys and programs, makes one bsg_probe_multi call and hands every
// member its own candidate mask.  Host-side C++ over the public C ABI only; callers' buffers are only
// read while their owners are blocked inside bsg_batcher_probe (cgo rule: nothing is retained).
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bloomgpu.h"

extern "C" int bsg_set_last_error_internal(int code, const char* msg);

namespace {
struct Req {
    const uint8_t* keys;
    const uint64_t* key_off;
    uint32_t n_keys;
    const uint8_t* kinds;
    const bsg_expr_op* prog;
    uint32_t prog_len;
    uint64_t* out_mask;
    int rc = BSG_OK;
    bool done = false;
    std::string err;
};
struct Batch {
    std::vector<Req*> reqs;
    uint32_t n_keys = 0;
    bool closed = false;   // the leader took it: no more members
    std::condition_variable done_cv;   // members sleep here: a finished launch wakes its own members only
};
}  // namespace

struct bsg_batcher {
    bsg_ctx* ctx = nullptr;
    const bsg_corpus* corpus = nullptr;
    uint32_t max_keys = 1024, max_queries = 256, window_us = 0;
    std::mutex mu;
    std::condition_variable cv;
    std::shared_ptr<Batch> open;   // the batch that accepts members (nullptr: none yet)
    uint32_t running = 0;          // launches in flight
    uint32_t last_size = 1;        // members of the previous batch: how much company a leader may expect
    uint64_t n_calls = 0, n_launches = 0, n_bypass = 0, max_batch = 0;
    // leader scratch (one launch at a time touches it: guarded by `running`)
    std::vector<uint8_t> keys, kinds;
    std::vector<uint64_t> key_off;
    std::vector<uint32_t> qbegin, pbegin;
    std::vector<bsg_expr_op> progs;
    std::vector<uint64_t> masks;
};

extern "C" int bsg_batcher_create(bsg_ctx* ctx, const bsg_corpus* corpus, uint32_t max_keys, uint32_t max_queries,
                                  uint32_t window_us, bsg_batcher** out) {
    if (!ctx || !corpus || !out) return bsg_set_last_error_internal(BSG_ERR_INVALID, "NULL argument");
    *out = nullptr;
    bsg_batcher* b = new (std::nothrow) bsg_batcher();
    if (!b) return bsg_set_last_error_internal(BSG_ERR_NOMEM, "batcher alloc");
    b->ctx = ctx;
    b->corpus = corpus;
    if (max_keys) b->max_keys = max_keys;
    if (max_queries) b->max_queries = max_queries < 65535u ? max_queries : 65535u;
    b->window_us = window_us;
    *out = b;
    return BSG_OK;
}

extern "C" void bsg_batcher_destroy(bsg_batcher* b) {
    if (!b) return;
    {   // callers must have returned; wait out a launch that is still unwinding
        std::unique_lock<std::mutex> lk(b->mu);
        b->cv.wait(lk, [&] { return b->running == 0; });
    }
    delete b;
}

extern "C" int bsg_batcher_stats(bsg_batcher* b, uint64_t* calls, uint64_t* launches, uint64_t* bypassed,
                                 uint64_t* largest_batch) {
    if (!b) return bsg_set_last_error_internal(BSG_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(b->mu);
    if (calls) *calls = b->n_calls;
    if (launches) *launches = b->n_launches;
    if (bypassed) *bypassed = b->n_bypass;
    if (largest_batch) *largest_batch = b->max_batch;
    return BSG_OK;
}

// Leader: one bsg_probe_multi call for the whole batch (mu NOT held).
static void run_batch(bsg_batcher* b, Batch& batch) {
    const uint64_t n_units = bsg_corpus_units(b->corpus);
    const size_t mask_words = (n_units + 63) / 64;
    const size_t nq = batch.reqs.size();
    b->keys.clear(); b->kinds.clear(); b->key_off.assign(1, 0); b->qbegin.assign(1, 0); b->pbegin.assign(1, 0);
    b->progs.clear();
    for (Req* r : batch.reqs) {
        const uint64_t base = b->key_off.back();
        const uint64_t nbytes = r->n_keys ? r->key_off[r->n_keys] - r->key_off[0] : 0;
        if (nbytes) b->keys.insert(b->keys.end(), r->keys + r->key_off[0], r->keys + r->key_off[0] + nbytes);
        for (uint32_t i = 0; i < r->n_keys; ++i) b->key_off.push_back(base + (r->key_off[i + 1] - r->key_off[0]));
        if (r->n_keys) b->kinds.insert(b->kinds.end(), r->kinds, r->kinds + r->n_keys);
        if (r->prog_len) b->progs.insert(b->progs.end(), r->prog, r->prog + r->prog_len);
        b->qbegin.push_back(b->qbegin.back() + r->n_keys);
        b->pbegin.push_back(b->pbegin.back() + r->prog_len);
    }
    b->masks.resize(nq * (mask_words ? mask_words : 1));
    uint8_t dummy = 0;
    int rc = bsg_probe_multi(b->ctx, b->corpus, b->keys.empty() ? &dummy : b->keys.data(), b->key_off.data(),
                             b->qbegin.back(), b->kinds.empty() ? &dummy : b->kinds.data(), static_cast<uint32_t>(nq),
                             b->qbegin.data(), b->progs.data(), b->pbegin.data(), b->masks.data());
    if (rc == BSG_OK) {
        for (size_t j = 0; j < nq; ++j)
            if (mask_words) memcpy(batch.reqs[j]->out_mask, b->masks.data() + j * mask_words, mask_words * 8);
        return;
    }
    if (nq == 1) {
        batch.reqs[0]->rc = rc;
        batch.reqs[0]->err = bsg_last_error();
        return;
    }
    // one member's bad program must not fail its neighbours: fall back to one call per member
    for (Req* r : batch.reqs) {
        r->rc = bsg_probe(b->ctx, b->corpus, r->keys, r->key_off, r->n_keys, r->kinds, r->prog, r->prog_len, nullptr,
                          r->out_mask);
        if (r->rc != BSG_OK) r->err = bsg_last_error();
    }
}

extern "C" int bsg_batcher_probe(bsg_batcher* b, const uint8_t* keys, const uint64_t* key_off, uint32_t n_keys,
                                 const uint8_t* key_kind, const bsg_expr_op* prog, uint32_t prog_len,
                                 uint64_t* out_mask) {
    if (!b || !out_mask || (n_keys && (!keys || !key_off || !key_kind)) || (prog_len && !prog))
        return bsg_set_last_error_internal(BSG_ERR_INVALID, "NULL argument");
    if (n_keys > b->max_keys) {  // a batch of its own anyway
        {
            std::lock_guard<std::mutex> lk(b->mu);
            ++b->n_calls;
            ++b->n_bypass;
        }
        return bsg_probe(b->ctx, b->corpus, keys, key_off, n_keys, key_kind, prog, prog_len, nullptr, out_mask);
    }
    Req req;
    req.keys = keys; req.key_off = key_off; req.n_keys = n_keys; req.kinds = key_kind;
    req.prog = prog; req.prog_len = prog_len; req.out_mask = out_mask;
    std::unique_lock<std::mutex> lk(b->mu);
    ++b->n_calls;
    if (!b->open || b->open->closed || b->open->n_keys + n_keys > b->max_keys || b->open->reqs.size() >= b->max_queries)
        b->open = std::make_shared<Batch>();
    std::shared_ptr<Batch> mine = b->open;
    const bool leader = mine->reqs.empty();
    mine->reqs.push_back(&req);
    mine->n_keys += n_keys;
    if (!leader) {
        if (b->window_us && (mine->n_keys >= b->max_keys || mine->reqs.size() >= b->max_queries)) b->cv.notify_all();  // full: wake the leader
        mine->done_cv.wait(lk, [&] { return req.done; });
    } else {
        // group commit: launch when nothing is in flight; an optional window lets an idle system collect members
        auto full = [&] { return mine->n_keys >= b->max_keys || mine->reqs.size() >= b->max_queries; };
        if (b->window_us && !full())
            b->cv.wait_for(lk, std::chrono::microseconds(b->window_us), full);
        b->cv.wait(lk, [&] { return b->running == 0; });
        // The members of the batch that just finished come back within microseconds of each other; the first one
        // back would otherwise launch alone and make the rest wait a whole launch.  If the previous batch had
        // company, give as many callers a few microseconds to arrive (bounded spin, lock released).
        if (b->last_size > 1 && !full()) {
            const uint32_t want = (b->last_size + 1) / 2;
            const auto t0 = std::chrono::steady_clock::now();
            while (mine->reqs.size() < want && !full() && b->running == 0 &&
                   std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(25)) {
                lk.unlock();
                std::this_thread::yield();
                lk.lock();
            }
            b->cv.wait(lk, [&] { return b->running == 0; });   // another leader may have slipped in meanwhile
        }
        mine->closed = true;
        if (b->open == mine) b->open.reset();
        ++b->running;
        ++b->n_launches;
        b->last_size = static_cast<uint32_t>(mine->reqs.size());
        if (mine->reqs.size() > b->max_batch) b->max_batch = mine->reqs.size();
        lk.unlock();
        run_batch(b, *mine);
        lk.lock();
        --b->running;
        for (Req* r : mine->reqs) r->done = true;
        mine->done_cv.notify_all();   // this batch's members
        b->cv.notify_all();           // leaders waiting for the device (and bsg_batcher_destroy)
    }
    lk.unlock();
    if (req.rc != BSG_OK) return bsg_set_last_error_internal(req.rc, req.err.c_str());
    return BSG_OK;
}

// ---- measurement plumbing (bench.py; not part of include/bloomgpu.h) ------------------------------------------
// n_threads host threads issue the same small query calls_each times, either straight through bsg_probe (batcher ==
// NULL) or through the batcher: the concurrency a Go host's goroutines would produce, without an interpreter
// between the calls.  *seconds = wall time from the common start to the last thread's return; out_mask (one mask)
// receives thread 0's last result for checking.
#include <atomic>
#include <thread>
extern "C" int bsg_debug_query_callers(bsg_ctx* ctx, const bsg_corpus* corpus, bsg_batcher* batcher, uint32_t n_threads,
                                       uint32_t calls_each, const uint8_t* keys, const uint64_t* key_off, uint32_t n_keys,
                                       const uint8_t* kinds, const bsg_expr_op* prog, uint32_t prog_len,
                                       uint64_t* out_mask, double* seconds) {
    if (!ctx || !corpus || !out_mask || !seconds || n_threads == 0)
        return bsg_set_last_error_internal(BSG_ERR_INVALID, "NULL argument");
    const size_t mask_words = (bsg_corpus_units(corpus) + 63) / 64;
    std::vector<std::vector<uint64_t>> masks(n_threads, std::vector<uint64_t>(mask_words ? mask_words : 1));
    std::atomic<uint32_t> ready{0};
    std::atomic<bool> go{false};
    std::atomic<int> first_rc{BSG_OK};
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; ++t)
        th.emplace_back([&, t] {
            ready.fetch_add(1);
            while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
            for (uint32_t i = 0; i < calls_each; ++i) {
                const int rc = batcher ? bsg_batcher_probe(batcher, keys, key_off, n_keys, kinds, prog, prog_len, masks[t].data())
                                       : bsg_probe(ctx, corpus, keys, key_off, n_keys, kinds, prog, prog_len, nullptr, masks[t].data());
                if (rc != BSG_OK) { int exp = BSG_OK; first_rc.compare_exchange_strong(exp, rc); return; }
            }
        });
    while (ready.load() < n_threads) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    for (auto& x : th) x.join();
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    memcpy(out_mask, masks[0].data(), (mask_words ? mask_words : 1) * 8);
    if (first_rc.load() != BSG_OK) return bsg_set_last_error_internal(first_rc.load(), "a caller thread failed");
    return BSG_OK;
}

// Same idea for the batch probe of the headline workload: n_threads host threads call bsg_probe(matrix out, no mask)
// calls_each times, cycling over n_corpora resident replicas (so no call finds the filters in L2), each into its own
// result buffer — pinned caller memory from bsg_host_alloc (pinned != 0) or plain pageable memory.  out_matrix
// receives thread 0's last result.
extern "C" int bsg_debug_probe_callers(bsg_ctx* ctx, const bsg_corpus* const* corpora, uint32_t n_corpora,
                                       uint32_t n_threads, uint32_t calls_each, const uint8_t* keys,
                                       const uint64_t* key_off, uint32_t n_keys, const uint8_t* kinds, int pinned,
                                       uint64_t* out_matrix, double* seconds) {
    if (!ctx || !corpora || n_corpora == 0 || !out_matrix || !seconds || n_threads == 0)
        return bsg_set_last_error_internal(BSG_ERR_INVALID, "NULL argument");
    const size_t words = static_cast<size_t>(bsg_corpus_units(corpora[0])) * ((n_keys + 63) / 64);
    std::vector<uint64_t*> bufs(n_threads, nullptr);
    std::vector<std::vector<uint64_t>> pageable;
    if (pinned) {
        for (uint32_t t = 0; t < n_threads; ++t) {
            void* p = nullptr;
            const int rc = bsg_host_alloc(ctx, (words ? words : 1) * 8, &p);
            if (rc != BSG_OK) { for (uint64_t* b : bufs) bsg_host_free(ctx, b); return rc; }
            bufs[t] = static_cast<uint64_t*>(p);
        }
    } else {
        pageable.assign(n_threads, std::vector<uint64_t>(words ? words : 1));
        for (uint32_t t = 0; t < n_threads; ++t) bufs[t] = pageable[t].data();
    }
    std::atomic<uint32_t> ready{0};
    std::atomic<bool> go{false};
    std::atomic<int> first_rc{BSG_OK};
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; ++t)
        th.emplace_back([&, t] {
            ready.fetch_add(1);
            while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
            for (uint32_t i = 0; i < calls_each; ++i) {
                const int rc = bsg_probe(ctx, corpora[(t + i) % n_corpora], keys, key_off, n_keys, kinds, nullptr, 0, bufs[t], nullptr);
                if (rc != BSG_OK) { int exp = BSG_OK; first_rc.compare_exchange_strong(exp, rc); return; }
            }
        });
    while (ready.load() < n_threads) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    for (auto& x : th) x.join();
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    memcpy(out_matrix, bufs[0], words * 8);
    if (pinned) for (uint64_t* b : bufs) bsg_host_free(ctx, b);
    if (first_rc.load() != BSG_OK) return bsg_set_last_error_internal(first_rc.load(), "a caller thread failed");
    return BSG_OK;
}
