/*
 * bloomgpu.h — C ABI of libbloomgpu.so: the B200 (sm_100a) implementation of
 * bloomsearch's bloom-filter build / probe hot path.
 *
 * This is the drop-in boundary a cgo shim binds (see INTEGRATION.md and
 * go/bloomgpu/).  The reference (danthegoodman1/bloomsearch @ 10735cf9, pure Go)
 * has no FFI seam of its own; the seam is the *bloom.BloomFilter value used at
 * exactly these call sites, each of which one entry point below replaces:
 *
 *   build   ingest.go:127-145  buildFilters / buildSizedBloomFilter
 *           (called from flush.go:204,253 and merge.go:516,771)   -> bsg_build
 *   probe   query_exec.go:75-159 evaluateBloom{Filters,Expression,Condition}
 *           (called from query_exec.go:399-404 file level,
 *            query_exec.go:592-597 block level)                    -> bsg_probe
 *   decode  file_format.go:392-448 parseFilterSection + bloom ReadFrom
 *           (per block per query in the reference)                 -> bsg_corpus_load[_sections]
 *   sizing  bloom.NewWithEstimates via ingest.go:140               -> bsg_estimate (helper)
 *
 * Conventions: every function returns 0 on success or a negative bsg_status;
 * bsg_strerror(code) gives a static string, bsg_last_error(ctx) the detail of
 * the last failure on the calling thread.  All pointers are plain host memory
 * owned by the caller and are not retained after the call returns (cgo rule).
 * No CPU fallback exists: without a CUDA device every entry point fails with
 * BSG_ERR_CUDA.  Entry points are thread-safe; concurrent probes on one ctx
 * run on separate streams.
 *
 * Bit-exact contract (what "same filter" means): bit b of a filter lives in
 * native-endian uint64 word b>>6 at mask 1<<(b&63) (bitset.BitSet), locations
 * are location(h,i) % m with h = the four MurmurHash3_x64_128 base hashes of
 * bloom/v3 (see oracle/bloomref.h for the restated algorithm).
 */
#ifndef BLOOMGPU_H
#define BLOOMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSG_ABI_VERSION 2

typedef enum bsg_status {
    BSG_OK = 0,
    BSG_ERR_INVALID = -1,    /* bad argument (NULL, m too large, malformed program, ...) */
    BSG_ERR_CUDA = -2,       /* CUDA runtime / driver failure or no device */
    BSG_ERR_NOMEM = -3,      /* host or device allocation failed */
    BSG_ERR_FORMAT = -4,     /* filter section framing / CRC error (per-unit detail in status array) */
    BSG_ERR_UNSUPPORTED = -5,/* feature not available (e.g. NCCL not loadable) */
    BSG_ERR_COMM = -6        /* NCCL failure */
} bsg_status;

typedef struct bsg_ctx bsg_ctx;
typedef struct bsg_corpus bsg_corpus;
typedef struct bsg_query bsg_query;
typedef struct bsg_keyset bsg_keyset;
typedef struct bsg_cache bsg_cache;

/* One bloom filter: m bits, k hash functions, words start at word_off (in uint64
 * units) inside the accompanying words array.  m == 0 means "filter absent"
 * (Go nil): an absent filter cannot disqualify (query_exec.go:137-151). */
typedef struct bsg_filter_desc {
    uint64_t m;
    uint64_t k;
    uint64_t word_off;
} bsg_filter_desc;

/* Condition kinds, query.go:478-484.  Kind selects which of a unit's three
 * filters a key is tested against. */
enum { BSG_KIND_FIELD = 0, BSG_KIND_TOKEN = 1, BSG_KIND_FIELDTOKEN = 2 };

/* Postfix form of a BloomExpression tree (query.go:499-503).  Leaf i is query
 * key i.  AND/OR pop `arg` values; AND 0 = true, OR 0 = false
 * (query_exec.go:105-123).  TRUE encodes a nil Condition / nil child, FALSE an
 * unknown expression or condition type (query_exec.go:121-123,155-156). */
enum { BSG_OP_LEAF = 0, BSG_OP_AND = 1, BSG_OP_OR = 2, BSG_OP_TRUE = 3, BSG_OP_FALSE = 4 };
typedef struct bsg_expr_op {
    uint32_t op;
    uint32_t arg;
} bsg_expr_op;
#define BSG_MAX_STACK 64 /* maximum evaluation-stack depth of a program */

/* ---- context ------------------------------------------------------------ */
int bsg_abi_version(void);
const char *bsg_strerror(int code);
const char *bsg_last_error(void); /* thread-local detail string */
int bsg_create(int device, bsg_ctx **out);
void bsg_destroy(bsg_ctx *ctx);
/* Run all work of this ctx's *resident* entry points on an existing CUDA stream
 * (cudaStream_t passed as void*); NULL restores the ctx's own stream. */
int bsg_set_stream(bsg_ctx *ctx, void *cuda_stream);
int bsg_synchronize(bsg_ctx *ctx);
/* Device properties the host needs for planning / reporting. */
int bsg_device_info(bsg_ctx *ctx, int *sm_count, size_t *smem_per_block_optin,
                    size_t *total_mem, int *cc_major, int *cc_minor);

/* Pinned, device-mapped host memory for RESULT buffers (optional).  When out_matrix of bsg_probe() lies in
 * such a buffer the probe kernel writes the rows straight into it and the call ends without a staging
 * copy; any other host memory (e.g. the Go heap) works too, through a pinned staging block.  Zero-filled. */
int bsg_host_alloc(bsg_ctx *ctx, size_t bytes, void **out);
int bsg_host_free(bsg_ctx *ctx, void *ptr);

/* ---- sizing helper (bloom.EstimateParameters + New's clamp) -------------- */
void bsg_estimate(uint64_t n, double fpr, uint64_t *m, uint64_t *k);

/* ---- K1: base hashes ----------------------------------------------------- *
 * keys packed back to back; key i = keys[key_off[i] .. key_off[i+1]).
 * out_hashes[4*i .. 4*i+3] = bloom/v3 baseHashes(key i).  Replaces the hashing
 * inside every AddString/TestString (ingest.go:142, query_exec.go:141-154). */
int bsg_hash_keys(bsg_ctx *ctx, const uint8_t *keys, const uint64_t *key_off, uint64_t n_keys,
                  uint64_t *out_hashes);

/* ---- K2/K3: build --------------------------------------------------------- *
 * Builds n_filters filters in one call.  Keys are grouped: group g owns keys
 * [group_begin[g], group_begin[g+1]) and inserts each into filter
 * group_filter[g] (its block-level filter, ingest.go:139-145) and — when
 * group_filter2 != NULL and group_filter2[g] != BSG_NO_FILTER — also into filter
 * group_filter2[g] (the file-level union filter, flush.go:221,253; duplicates
 * across groups are harmless because insertion is an idempotent OR, only the
 * (m,k) sizing needs the exact union count, which the host's entry sets give).
 * Each key is hashed once for both.  Several groups may share a filter.
 * desc[f].word_off locates filter f inside out_words (n_words uint64, native
 * endian), which the callee zero-initialises. */
#define BSG_NO_FILTER 0xFFFFFFFFu
int bsg_build(bsg_ctx *ctx, const uint8_t *keys, const uint64_t *key_off, uint64_t n_keys,
              const uint64_t *group_begin, uint32_t n_groups, const uint32_t *group_filter,
              const uint32_t *group_filter2, const bsg_filter_desc *desc, uint32_t n_filters,
              uint64_t *out_words, uint64_t n_words);

/* Fused field::token build (ingest.go:95-102 addFieldToken + tokenizer.go:508-511): same contract
 * as bsg_build, but entry i of a group is the pair (pair_path[i], pair_token[i]) of indexes into ONE
 * string table (strings / str_off, n_strings entries) and its key is
 * strings[path] + "::" + strings[token], hashed on the device as a byte stream — the host never
 * materialises (or ships over PCIe) the joined keys. */
int bsg_build_fieldtokens(bsg_ctx *ctx, const uint8_t *strings, const uint64_t *str_off, uint64_t n_strings,
                          const uint32_t *pair_path, const uint32_t *pair_token, uint64_t n_pairs,
                          const uint64_t *group_begin, uint32_t n_groups, const uint32_t *group_filter,
                          const uint32_t *group_filter2, const bsg_filter_desc *desc, uint32_t n_filters,
                          uint64_t *out_words, uint64_t n_words);

/* Exact distinct counts (what the Go maps of bloomEntrySets provide, ingest.go:24-45,105-123):
 * keys may repeat inside and across groups.  out_group_counts[g] = number of distinct keys of
 * group g; if group_parent != NULL, out_parent_counts[p] = number of distinct keys of the union
 * of all groups with group_parent[g] == p (the file-level union count, flush.go:221,253).
 * These are the n of NewWithEstimates(max(n,1), fpr) (ingest.go:139-140).  Exact: emissions are
 * ordered by (group, 64-bit hash) with a radix sort and neighbours with equal hashes are compared
 * byte for byte (a hash collision between distinct keys is counted as two keys, and triggers a
 * second pass ordered by 128 hash bits so that their repeats cannot interleave). */
int bsg_count_distinct(bsg_ctx *ctx, const uint8_t *keys, const uint64_t *key_off, uint64_t n_keys,
                       const uint64_t *group_begin, uint32_t n_groups, const uint32_t *group_parent,
                       uint32_t n_parents, uint64_t *out_group_counts, uint64_t *out_parent_counts);

/* ---- resident key sets (build side) ----------------------------------------- *
 * A key set is a batch of grouped keys resident in HBM, the build side's counterpart of bsg_query:
 * the emissions of a flush cross PCIe once, then exact distinct counts (the `n` of
 * NewWithEstimates, ingest.go:139-140), the sizing on the host and the build itself all read the
 * resident copy.  Groups must cover [0, n_keys) (CSR group_begin).  bsg_keyset_build is
 * asynchronous on the ctx stream; d_out_words is DEVICE memory of n_words uint64 (zeroed by the
 * callee) or NULL for an internal buffer — e.g. symmetric memory from bsg_comm_alloc, so that
 * partial file-level filters can be OR-combined across GPUs without leaving the device. */
int bsg_keyset_create(bsg_ctx *ctx, const uint8_t *keys, const uint64_t *key_off, uint64_t n_keys,
                      const uint64_t *group_begin, uint32_t n_groups, bsg_keyset **out);
int bsg_keyset_count_distinct(bsg_ctx *ctx, bsg_keyset *ks, const uint32_t *group_parent, uint32_t n_parents,
                              uint64_t *out_group_counts, uint64_t *out_parent_counts);
int bsg_keyset_set_filters(bsg_ctx *ctx, bsg_keyset *ks, const uint32_t *group_filter,
                           const uint32_t *group_filter2, const bsg_filter_desc *desc, uint32_t n_filters,
                           uint64_t n_words);
int bsg_keyset_build(bsg_ctx *ctx, bsg_keyset *ks, uint64_t *d_out_words);
int bsg_keyset_fetch(bsg_ctx *ctx, bsg_keyset *ks, uint64_t *out_words); /* synchronises, copies the last build */
const uint64_t *bsg_keyset_device_words(const bsg_keyset *ks);           /* where the last build wrote (device) */
void bsg_keyset_free(bsg_keyset *ks);

/* ---- corpus residency ------------------------------------------------------ *
 * A corpus is n_units "units" (data blocks, or files for the file-level stage),
 * each with up to three filters: desc[3*u + kind].  The words are copied to HBM
 * and re-laid-out (16-byte aligned, unit-contiguous) for the probe kernels.
 * big_endian != 0: `words` holds the on-disk big-endian uint64 words of
 * bitset.WriteTo and is byte-swapped on the device. */
int bsg_corpus_load(bsg_ctx *ctx, const bsg_filter_desc *desc, uint64_t n_units,
                    const uint64_t *words, uint64_t n_words, int big_endian, bsg_corpus **out);
/* Load straight from raw filter sections (file_format.go:343-385 framing):
 * unit u = sections[sec_off[u] .. sec_off[u+1]).  Framing is parsed and the
 * CRC32C verified (verify_crc != 0) — on the device; big-endian words are
 * swapped on the device.  unit_status (nullable, n_units ints) receives 0 or
 * the per-unit BSG_ERR_FORMAT detail code.  Per-block error isolation as in the
 * reference (query_exec.go:580-590: the error is recorded, the loop continues
 * and the block is NOT scanned): a unit that fails to parse never survives —
 * its bit in every candidate mask of this corpus is 0 — and the caller reports
 * unit_status[u] as that block's error.  (Its matrix row reads all ones: it has
 * no filter that could disqualify a key; the mask is authoritative.)  Returns
 * BSG_OK even if some units failed; *n_bad (nullable) counts them. */
int bsg_corpus_load_sections(bsg_ctx *ctx, const uint8_t *sections, const uint64_t *sec_off,
                             uint64_t n_units, int verify_crc, int32_t *unit_status,
                             uint64_t *n_bad, bsg_corpus **out);
void bsg_corpus_free(bsg_corpus *corpus);
uint64_t bsg_corpus_units(const bsg_corpus *corpus);
/* Bytes of filter bitsets resident in HBM (Σ_b S_b of SURVEY.md §8d), optionally
 * only for the kinds in kind_mask (bit kind). */
uint64_t bsg_corpus_bitset_bytes(const bsg_corpus *corpus, uint32_t kind_mask);
/* HBM held by the corpus (bitsets + descriptor tables): the resident cache's unit of account. */
uint64_t bsg_corpus_device_bytes(const bsg_corpus *corpus);
/* Copy unit u's descriptors (3) and words back out (tests / round trips). */
int bsg_corpus_unit_desc(const bsg_corpus *corpus, uint64_t unit, bsg_filter_desc out_desc[3]);

/* ---- resident filter cache -------------------------------- *
 * The reference decodes a file's filters per query and drops them (query_exec.go:399-412, :572-615); the
 * GPU path keeps them in HBM.  Corpora are keyed by a caller-chosen file id (one id space per cache: use
 * one cache for file-level corpora and one for block-level corpora, or distinct ids), held under a byte
 * budget with least-recently-used eviction, and pinned while in use:
 *   acquire   hit: pins and returns the corpus; miss: *out = NULL (the caller reads the sections and inserts)
 *   insert*   takes ownership of / loads a corpus for file_id (replacing an older one), evicts unpinned
 *             entries down to the budget, returns it pinned when out != NULL
 *   release   unpins; an entry that was invalidated or evicted while pinned is freed on its last release
 *   invalidate  the file was replaced by a merge (merge.go:529-536) or tombstoned: never served again
 * Thread-safe.  A corpus obtained here must not be passed to bsg_corpus_free. */
int bsg_cache_create(bsg_ctx *ctx, uint64_t budget_bytes, bsg_cache **out);
void bsg_cache_destroy(bsg_cache *cache);
int bsg_cache_acquire(bsg_cache *cache, uint64_t file_id, const bsg_corpus **out);
int bsg_cache_insert(bsg_cache *cache, uint64_t file_id, bsg_corpus *corpus, const bsg_corpus **out);
int bsg_cache_insert_sections(bsg_cache *cache, uint64_t file_id, const uint8_t *sections, const uint64_t *sec_off,
                              uint64_t n_units, int verify_crc, int32_t *unit_status, uint64_t *n_bad,
                              const bsg_corpus **out);
void bsg_cache_release(bsg_cache *cache, const bsg_corpus *corpus);
int bsg_cache_invalidate(bsg_cache *cache, uint64_t file_id);
int bsg_cache_stats(bsg_cache *cache, uint64_t *used_bytes, uint64_t *entries, uint64_t *hits, uint64_t *misses,
                    uint64_t *evictions, uint64_t *invalidations);

/* ---- K4/K5: probe ----------------------------------------------------------- *
 * Tests n_keys keys against every unit of the corpus.
 *   key_kind[q]   BSG_KIND_* : which filter of the unit key q is tested on
 *   out_matrix    nullable; n_units rows of ceil(n_keys/64) uint64 words,
 *                 bit q of row u = TestString(key q) on unit u (absent filter => 1)
 *   prog/prog_len nullable postfix BloomExpression; NULL/0 => every unit survives
 *                 (query_exec.go:81-83)
 *   out_mask      nullable; ceil(n_units/64) uint64 words, bit u = unit u survives
 * FieldToken keys are passed already joined as field + "::" + token
 * (tokenizer.go:508-511); bytes are used verbatim (no normalisation).
 * Thread-safe and re-entrant (one pooled stream + scratch per call); the caller's buffers are plain
 * host memory and are not retained.  Per call: one host-to-device copy of the packed batch, one
 * kernel when the corpus is staged (hashing is fused into the probe), and — matrix without mask, up
 * to 8 MB — the rows arrive in pinned host memory while the kernel runs (no copy back). */
int bsg_probe(bsg_ctx *ctx, const bsg_corpus *corpus, const uint8_t *keys, const uint64_t *key_off,
              uint32_t n_keys, const uint8_t *key_kind, const bsg_expr_op *prog, uint32_t prog_len,
              uint64_t *out_matrix, uint64_t *out_mask);

/* Several queries in one call.  The keys of all queries are packed as for bsg_probe;
 * query j owns keys [query_key_begin[j], query_key_begin[j+1]) and the postfix program
 * progs[prog_begin[j] .. prog_begin[j+1]), whose LEAF arguments index ITS OWN keys (0-based); an empty
 * program keeps every unit (query_exec.go:81-83).  The corpus is probed ONCE for the union of the keys (one
 * pass per 1 024 keys: the staged kernels stream every filter byte once per pass whatever the number of
 * keys), then every query's expression is evaluated on its own columns.  out_masks: n_queries rows of
 * ceil(n_units/64) words, bit u of row j = unit u survives query j — identical to n_queries separate
 * bsg_probe calls.  Replaces n_queries executions of query_exec.go:572-615 over the same blocks. */
int bsg_probe_multi(bsg_ctx *ctx, const bsg_corpus *corpus, const uint8_t *keys, const uint64_t *key_off,
                    uint32_t n_keys, const uint8_t *key_kind, uint32_t n_queries,
                    const uint32_t *query_key_begin, const bsg_expr_op *progs, const uint32_t *prog_begin,
                    uint64_t *out_masks);

/* Query batcher: merges CONCURRENT bsg_batcher_probe callers (one goroutine per query in the reference,
 * query_exec.go:201-433) into bsg_probe_multi launches.  Group commit: a caller that finds no launch in
 * flight launches at once; callers that arrive while one is running form the next batch (closed at max_keys
 * keys / max_queries queries, 0 = defaults 1024 / 256; later arrivals open another), launched when the running
 * one ends.  window_us > 0 additionally lets the first caller of a batch wait that long for company.  Blocking and
 * thread-safe; each caller gets exactly the mask bsg_probe would have returned; a member whose program is
 * malformed fails alone.  The corpus must outlive the batcher (pin it in the cache). */
typedef struct bsg_batcher bsg_batcher;
int bsg_batcher_create(bsg_ctx *ctx, const bsg_corpus *corpus, uint32_t max_keys, uint32_t max_queries,
                       uint32_t window_us, bsg_batcher **out);
void bsg_batcher_destroy(bsg_batcher *batcher);
int bsg_batcher_probe(bsg_batcher *batcher, const uint8_t *keys, const uint64_t *key_off, uint32_t n_keys,
                      const uint8_t *key_kind, const bsg_expr_op *prog, uint32_t prog_len, uint64_t *out_mask);
int bsg_batcher_stats(bsg_batcher *batcher, uint64_t *calls, uint64_t *launches, uint64_t *bypassed,
                      uint64_t *largest_batch);

/* Hierarchical probe — the reference's two stages in one call: file-level filters first
 * (query_exec.go:399-406), then block-level filters only for blocks whose file survived
 * (query_exec.go:572-615); the surviving units are compacted on the device between the stages.
 * bsg_corpus_set_parents records, for every unit of `blocks`, the index of its file in `files`.
 * out_file_mask (nullable): ceil(files/64) words; out_block_mask: ceil(blocks/64) words,
 * bit u = block u survives (its file survived AND its own filters pass the expression). */
int bsg_corpus_set_parents(bsg_ctx *ctx, bsg_corpus *corpus, const uint32_t *parent, uint64_t n_units,
                           uint64_t n_parent_units);
int bsg_probe_hierarchical(bsg_ctx *ctx, const bsg_corpus *files, const bsg_corpus *blocks,
                           const uint8_t *keys, const uint64_t *key_off, uint32_t n_keys,
                           const uint8_t *key_kind, const bsg_expr_op *prog, uint32_t prog_len,
                           uint64_t *out_file_mask, uint64_t *out_block_mask);

/* Resident form of the same call, for callers that keep a query on the device
 * and for measurement: create uploads + hashes the keys once; run launches the
 * probe (asynchronously, on the ctx stream); fetch copies results to the host. */
enum { BSG_PROBE_AUTO = 0, BSG_PROBE_STAGED = 1, BSG_PROBE_GATHER = 2 };
/* OR into `path`: run only the probe kernel(s) (bit matrix), skip the mask kernel. */
#define BSG_RUN_MATRIX_ONLY 0x100
int bsg_query_create(bsg_ctx *ctx, const bsg_corpus *corpus, const uint8_t *keys,
                     const uint64_t *key_off, uint32_t n_keys, const uint8_t *key_kind,
                     const bsg_expr_op *prog, uint32_t prog_len, bsg_query **out);
/* want_matrix == 0: the caller will fetch only the candidate mask.  A query of <= 32 keys on the gather path then
 * short-circuits like evaluateBloomExpression (query_exec.go:105-119): a unit's keys stop being tested once its
 * expression is decided, and the matrix rows are only an upper bound of the membership bits (the mask is exact). */
int bsg_query_run(bsg_ctx *ctx, const bsg_corpus *corpus, bsg_query *q, int path, int want_matrix);
int bsg_query_fetch(bsg_ctx *ctx, bsg_query *q, uint64_t n_units, uint64_t *out_matrix,
                    uint64_t *out_mask);
void bsg_query_free(bsg_query *q);
/* Second stage of a hierarchical probe with resident queries: run `q` on `blocks` only for units whose
 * parent (bsg_corpus_set_parents) survived `parent_q`'s last run on the files corpus; the candidate
 * mask stays on the device (bsg_query_device_mask: ceil(n_units/64) uint64 words, padded to a multiple
 * of 2 words) so that bsg_allgather_masks_device can exchange it without a host round trip. */
int bsg_query_run_child(bsg_ctx *ctx, const bsg_corpus *blocks, bsg_query *q, const bsg_query *parent_q, int path);
const uint64_t *bsg_query_device_mask(const bsg_query *q);
/* Number of kernels the last bsg_query_run launched (for launch accounting). */
int bsg_query_last_launches(const bsg_query *q);

/* ---- device-event timing on the ctx stream (measurement only) ------------- */
int bsg_timer_begin(bsg_ctx *ctx);
int bsg_timer_end(bsg_ctx *ctx, float *elapsed_ms); /* records, synchronises, returns ms */

/* ---- multi-GPU (one process per GPU) --------------------------------------- *
 * nccl_unique_id: the 128-byte ncclUniqueId created on rank 0 (bsg_comm_unique_id)
 * and distributed by the host (the Go side would ship it over its own RPC).
 * Environment, read at bsg_comm_init: BSG_COMM_P2P=0 forces the NCCL path; BSG_COMM_TIMEOUT_S (default
 * 120, 0 = wait for ever) bounds how long a peer-memory collective kernel waits for a peer's flag — a
 * peer that never arrives traps the kernel, and the next call on this ctx returns BSG_ERR_CUDA
 * instead of the GPU spinning for ever. */
int bsg_comm_unique_id(uint8_t out_id[128]);
int bsg_comm_init(bsg_ctx *ctx, int rank, int world, const uint8_t nccl_unique_id[128]);
/* rank / world of the communicator; *peer_memory = 1 when the collectives run as single kernels over
 * peer memory (CUDA IPC mappings, NVLink loads / stores), 0 when they use NCCL send/recv;
 * *last_nvlink_bytes = bytes this rank moved over NVLink in its last collective. */
int bsg_comm_info(bsg_ctx *ctx, int *rank, int *world, int *peer_memory, uint64_t *last_nvlink_bytes);
/* Symmetric device memory (collective: every rank calls with the same size, in the same order).  The
 * returned pointer is this rank's buffer; the library maps every peer's copy so that the device
 * collectives below can read and write peers directly.  Rounded up to 2 MiB. */
int bsg_comm_alloc(bsg_ctx *ctx, size_t bytes, void **out_dev);
int bsg_comm_free(bsg_ctx *ctx, void *dev);
/* In-place bitwise OR across ranks of equal-shape partial bitsets (file-level filters built from
 * disjoint shards of a file's entries: flush.go:221,253 builds that filter from the union of the entry
 * sets, and OR of partials of equal (m,k) is the same bitset).  Result identical on
 * every rank.  The _device form takes a 16-byte aligned pointer into bsg_comm_alloc memory, allocates
 * nothing, and is asynchronous on the ctx stream: one kernel per rank (rank r ORs slice r out of every
 * peer and stores it into every peer; 2*(W-1)/W * bytes over NVLink per rank).  The host form stages
 * through a cached symmetric buffer and synchronises. */
int bsg_or_reduce_device(bsg_ctx *ctx, uint64_t *d_words, uint64_t n_words);
int bsg_or_reduce(bsg_ctx *ctx, uint64_t *words, uint64_t n_words);
/* Gather every rank's candidate mask (n_words each) into all (world*n_words, rank-major).  _device:
 * d_local is any device memory, d_all is bsg_comm_alloc memory; asynchronous on the ctx stream.  A rank
 * must have consumed (stream-ordered) the previous result in d_all before it calls again. */
int bsg_allgather_masks_device(bsg_ctx *ctx, const uint64_t *d_local, uint64_t n_words, uint64_t *d_all);
int bsg_allgather_masks(bsg_ctx *ctx, const uint64_t *local, uint64_t n_words, uint64_t *all);
/* Sharded hierarchical probe, one collective host call per query: every rank runs
 * bsg_probe_hierarchical on its shard (files + their blocks), the per-rank block masks (padded to
 * mask_words each) are all-gathered on the device, and out_all_block_masks receives
 * world * mask_words words, rank-major (query_exec.go:372-433,572-615 with files dealt to GPUs). */
int bsg_probe_hierarchical_gather(bsg_ctx *ctx, const bsg_corpus *files, const bsg_corpus *blocks,
                                  const uint8_t *keys, const uint64_t *key_off, uint32_t n_keys,
                                  const uint8_t *key_kind, const bsg_expr_op *prog, uint32_t prog_len,
                                  uint64_t mask_words, uint64_t *out_all_block_masks);

#ifdef __cplusplus
}
#endif
#endif /* BLOOMGPU_H */
