// Resident filter cache.
//
// The reference decodes a file's filters for every query and drops them again (query_exec.go:399-412 at the
// file level, :572-615 via parseFilterSection at the block level).  The GPU path keeps them in HBM instead;
// this is the bookkeeping that makes that safe: corpora keyed by file id, a byte budget with
// least-recently-used eviction, pin counts so that a corpus in use by a running query is never freed under
// it, and explicit invalidation for the two events that retire a file — a merge committing its replacement
// (merge.go:529-536: MetaStore.Update adds the merged file and tombstones the sources) and
// DataStore.TombstoneFile.  Host-side C++ over the public C ABI only.
#include <cstdint>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/bloomgpu.h"

extern "C" int bsg_set_last_error_internal(int code, const char* msg);

namespace {
struct Entry {
    uint64_t file_id = 0;
    bsg_corpus* corpus = nullptr;
    uint64_t bytes = 0;
    uint64_t last_use = 0;
    uint32_t pins = 0;
    bool dead = false;   // invalidated or evicted while pinned: freed on the last release
};
}  // namespace

struct bsg_cache {
    bsg_ctx* ctx = nullptr;
    uint64_t budget = 0, used = 0, tick = 0;
    uint64_t hits = 0, misses = 0, evictions = 0, invalidations = 0;
    std::mutex mu;
    std::unordered_map<uint64_t, Entry*> live;                // file id -> current entry
    std::unordered_map<const bsg_corpus*, Entry*> by_corpus;  // every entry that still owns a corpus
};

static void drop(bsg_cache* c, Entry* e) {  // caller holds mu; e is unpinned
    c->by_corpus.erase(e->corpus);
    c->used -= e->bytes;
    bsg_corpus_free(e->corpus);
    delete e;
}

// retire an entry: out of the id map now, freed now or on its last release
static void retire(bsg_cache* c, Entry* e) {
    auto it = c->live.find(e->file_id);
    if (it != c->live.end() && it->second == e) c->live.erase(it);
    if (e->pins == 0) drop(c, e);
    else e->dead = true;
}

static void evict_to_budget(bsg_cache* c, const Entry* keep) {
    while (c->used > c->budget) {
        Entry* victim = nullptr;
        for (auto& kv : c->live) {
            Entry* e = kv.second;
            if (e == keep || e->pins) continue;
            if (!victim || e->last_use < victim->last_use) victim = e;
        }
        if (!victim) return;  // everything left is pinned: over budget until a release
        ++c->evictions;
        retire(c, victim);
    }
}

extern "C" int bsg_cache_create(bsg_ctx* ctx, uint64_t budget_bytes, bsg_cache** out) {
    if (!ctx || !out) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_cache_create: NULL argument");
    bsg_cache* c = new (std::nothrow) bsg_cache();
    if (!c) return bsg_set_last_error_internal(BSG_ERR_NOMEM, "cache alloc");
    c->ctx = ctx;
    c->budget = budget_bytes;
    *out = c;
    return BSG_OK;
}

extern "C" void bsg_cache_destroy(bsg_cache* c) {
    if (!c) return;
    for (auto& kv : c->by_corpus) {
        bsg_corpus_free(kv.second->corpus);
        delete kv.second;
    }
    delete c;
}

extern "C" int bsg_cache_acquire(bsg_cache* c, uint64_t file_id, const bsg_corpus** out) {
    if (!c || !out) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_cache_acquire: NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->live.find(file_id);
    if (it == c->live.end()) {
        ++c->misses;
        *out = nullptr;
        return BSG_OK;
    }
    ++c->hits;
    it->second->pins++;
    it->second->last_use = ++c->tick;
    *out = it->second->corpus;
    return BSG_OK;
}

extern "C" int bsg_cache_insert(bsg_cache* c, uint64_t file_id, bsg_corpus* corpus, const bsg_corpus** out) {
    if (!c || !corpus) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_cache_insert: NULL argument");
    Entry* e = new (std::nothrow) Entry();
    if (!e) return bsg_set_last_error_internal(BSG_ERR_NOMEM, "cache entry alloc");
    e->file_id = file_id;
    e->corpus = corpus;
    e->bytes = bsg_corpus_device_bytes(corpus);
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->live.find(file_id);
    if (it != c->live.end()) retire(c, it->second);   // a newer version of the file's filters replaces the old one
    e->last_use = ++c->tick;
    e->pins = out ? 1 : 0;
    c->live[file_id] = e;
    c->by_corpus[corpus] = e;
    c->used += e->bytes;
    evict_to_budget(c, e);
    if (out) *out = corpus;
    return BSG_OK;
}

extern "C" int bsg_cache_insert_sections(bsg_cache* c, uint64_t file_id, const uint8_t* sections, const uint64_t* sec_off,
                                         uint64_t n_units, int verify_crc, int32_t* unit_status, uint64_t* n_bad,
                                         const bsg_corpus** out) {
    if (!c) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_cache_insert_sections: NULL argument");
    bsg_corpus* corpus = nullptr;
    int rc = bsg_corpus_load_sections(c->ctx, sections, sec_off, n_units, verify_crc, unit_status, n_bad, &corpus);
    if (rc) return rc;
    rc = bsg_cache_insert(c, file_id, corpus, out);
    if (rc) bsg_corpus_free(corpus);
    return rc;
}

extern "C" void bsg_cache_release(bsg_cache* c, const bsg_corpus* corpus) {
    if (!c || !corpus) return;
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->by_corpus.find(corpus);
    if (it == c->by_corpus.end()) return;
    Entry* e = it->second;
    if (e->pins) e->pins--;
    if (e->pins == 0) {
        if (e->dead) drop(c, e);
        else evict_to_budget(c, nullptr);
    }
}

extern "C" int bsg_cache_invalidate(bsg_cache* c, uint64_t file_id) {
    if (!c) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_cache_invalidate: NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->live.find(file_id);
    if (it == c->live.end()) return BSG_OK;
    ++c->invalidations;
    retire(c, it->second);
    return BSG_OK;
}

extern "C" int bsg_cache_stats(bsg_cache* c, uint64_t* used_bytes, uint64_t* entries, uint64_t* hits, uint64_t* misses,
                               uint64_t* evictions, uint64_t* invalidations) {
    if (!c) return bsg_set_last_error_internal(BSG_ERR_INVALID, "bsg_cache_stats: NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    if (used_bytes) *used_bytes = c->used;
    if (entries) *entries = c->live.size();
    if (hits) *hits = c->hits;
    if (misses) *misses = c->misses;
    if (evictions) *evictions = c->evictions;
    if (invalidations) *invalidations = c->invalidations;
    return BSG_OK;
}
