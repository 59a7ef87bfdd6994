/*
 * bloomref — CPU ORACLE for the bloomsearch bloom build / probe hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under bloomsearch_b200/ (the product) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker and
 * as the timed CPU baseline — never as the thing shipped.
 *
 * It restates, in plain C, what the Go reference computes on this path:
 *   - reference call sites: ingest.go:127-145 (buildSizedBloomFilter),
 *     query_exec.go:75-159 (evaluateBloom{Filters,Expression,Condition}),
 *     tokenizer.go:508-511 (makeFieldTokenKey),
 *     file_format.go:343-448 (encodeFilterSection / parseFilterSection);
 *   - the arithmetic itself lives in third-party modules that are NOT vendored
 *     in /root/reference: github.com/bits-and-blooms/bloom/v3 v3.7.0 and
 *     github.com/bits-and-blooms/bitset v1.10.0 (go.mod:6,13).  Their published
 *     algorithm is restated here: MurmurHash3_x64_128(seed 0) of data and of
 *     data||0x01 -> 4 base hashes; location(h,i); % m; bitset word/bit order;
 *     big-endian WriteTo framing; EstimateParameters.
 *
 * PARITY STATUS: "parity unpinned" at the bit level.  No Go toolchain exists in
 * the build container and the reference's tests hold no golden bitsets
 * (SURVEY.md §8c).  What IS pinned: the murmur3 core — bref_base_hashes for
 * every key length and tail shape — against Austin Appleby's canonical
 * MurmurHash3.cpp, compiled unmodified out of scikit-learn's tree into
 * oracle/_ref (oracle/Makefile `ref`, oracle/murmur_canonical.py,
 * tests/test_oracle.py), and against the public MurmurHash3_x64_128 vectors
 * (the same ones spaolacci/murmur3's test-suite uses, which bloom/v3 documents
 * strict equivalence with); what is NOT pinned against third-party code is the
 * composition (location, % m, bit order, WriteTo, EstimateParameters).  Also
 * pinned: CRC32C against its
 * standard check value, the (m,k) table of SURVEY.md §8(c), the semantic pins of
 * the reference's tests (tree semantics, sizing, FPR budget, round trip), and
 * an independent Python restatement (oracle/bloomref.py) that must agree with
 * this file bit for bit.
 */
#ifndef BLOOMREF_H
#define BLOOMREF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- bloom/v3 murmur.go (restated) ------------------------------------- */
void bref_murmur3_x64_128(const void *data, size_t len, uint32_t seed, uint64_t out[2]);
/* bloom.go baseHashes: (h0,h1)=murmur(data), (h2,h3)=murmur(data||0x01), seed 0 */
void bref_base_hashes(const uint8_t *data, size_t len, uint64_t h[4]);
/* bloom.go location(): h[i%2] + i*h[2+(((i+(i%2))%4)/2)], u64 wraparound */
uint64_t bref_location(const uint64_t h[4], uint64_t i);
/* bloom.go EstimateParameters(n,p) (float64 formula), then New()'s clamp to >=1 */
void bref_estimate_parameters(uint64_t n, double p, uint64_t *m, uint64_t *k);

/* ---- bloom.BloomFilter + bitset.BitSet --------------------------------- */
typedef struct bref_filter {
    uint64_t m;       /* bits */
    uint64_t k;       /* hash functions */
    uint64_t nwords;  /* ceil(m/64) */
    uint64_t *words;  /* bit b -> words[b>>6] & (1<<(b&63)) */
} bref_filter;

bref_filter *bref_filter_new(uint64_t m, uint64_t k);           /* bloom.New */
bref_filter *bref_filter_new_with_estimates(uint64_t n, double fpr);
void bref_filter_free(bref_filter *f);
void bref_filter_add(bref_filter *f, const uint8_t *data, size_t len);      /* Add / AddString */
int bref_filter_test(const bref_filter *f, const uint8_t *data, size_t len); /* Test / TestString */
int bref_filter_equal(const bref_filter *a, const bref_filter *b);
/* accessors for ctypes */
uint64_t bref_filter_m(const bref_filter *f);
uint64_t bref_filter_k(const bref_filter *f);
uint64_t bref_filter_nwords(const bref_filter *f);
uint64_t *bref_filter_words(bref_filter *f);

/* WriteTo/ReadFrom: [u64 BE m][u64 BE k][u64 BE bitlen=m][ceil(m/64) x u64 BE] */
size_t bref_filter_serialized_size(const bref_filter *f);
size_t bref_filter_write_to(const bref_filter *f, uint8_t *out);
/* returns NULL on truncated / inconsistent input; *consumed = bytes read */
bref_filter *bref_filter_read_from(const uint8_t *in, size_t len, size_t *consumed);

/* ---- ingest.go:139-145 buildSizedBloomFilter --------------------------- */
/* keys packed: key i = bytes[key_off[i] .. key_off[i+1]) */
bref_filter *bref_build_sized_filter(const uint8_t *bytes, const uint64_t *key_off,
                                     uint64_t n_keys, double fpr);

/* ---- file_format.go:343-448 filter section codec ----------------------- */
uint32_t bref_crc32c(const uint8_t *data, size_t len);     /* SSE4.2 instruction where present (as Go does), else _sw */
uint32_t bref_crc32c_sw(const uint8_t *data, size_t len);  /* table-driven, slicing-by-8 */
/* filters[3] = field, token, fieldtoken; NULL entry = absent. returns bytes written
 * (call with out==NULL to size). */
size_t bref_section_encode(const bref_filter *const filters[3], uint8_t *out);
/* returns 0 ok; <0 error (-1 too small, -2 crc, -3 flags, -4 truncated len prefix,
 * -5 length exceeds remainder, -6 filter decode, -7 trailing bytes). On success
 * filters[i] is a fresh filter or NULL when absent. */
int bref_section_parse(const uint8_t *section, size_t len, bref_filter *filters[3]);

/* ---- query_exec.go:75-159 expression tree ------------------------------ */
enum { BREF_EXPR_CONDITION = 0, BREF_EXPR_AND = 1, BREF_EXPR_OR = 2, BREF_EXPR_UNKNOWN = 3 };
enum { BREF_COND_FIELD = 0, BREF_COND_TOKEN = 1, BREF_COND_FIELD_TOKEN = 2, BREF_COND_UNKNOWN = 3 };

typedef struct bref_expr {
    int32_t type;             /* BREF_EXPR_* */
    int32_t has_condition;    /* Condition != nil */
    int32_t cond_type;        /* BREF_COND_* */
    int32_t n_children;
    const uint8_t *field; uint64_t field_len;
    const uint8_t *token; uint64_t token_len;
    const struct bref_expr *children;
} bref_expr;

/* evaluateBloomFilters: expr == NULL <=> bloomQuery == nil || Expression == nil */
int bref_evaluate_bloom_filters(const bref_filter *field_f, const bref_filter *token_f,
                                const bref_filter *fieldtoken_f, const bref_expr *expr);

/* ---- flat (postfix) form of the same tree, shared with the C ABI ------- */
enum { BREF_OP_LEAF = 0, BREF_OP_AND = 1, BREF_OP_OR = 2, BREF_OP_TRUE = 3, BREF_OP_FALSE = 4 };
typedef struct bref_op { uint32_t op; uint32_t arg; } bref_op;
/* leaf_bits[i] = result of leaf i; returns tree value, or -1 on malformed program */
int bref_eval_postfix(const bref_op *prog, uint32_t prog_len, const uint8_t *leaf_bits,
                      uint32_t n_leaves);

/* ---- bulk helpers (same packed inputs the C ABI takes) ------------------ */
typedef struct bref_desc { uint64_t m, k, word_off; } bref_desc; /* m==0: absent */

/* Build n_filters filters.  Keys of group g are [group_begin[g], group_begin[g+1]);
 * each is inserted into filter group_filter[g] and, if group_filter2 != NULL and
 * group_filter2[g] != 0xFFFFFFFF, also into that one (the file-level union,
 * flush.go:221,253).  out_words must be zeroed, native-endian. */
void bref_build_filters(const uint8_t *bytes, const uint64_t *key_off,
                        const uint64_t *group_begin, uint32_t n_groups,
                        const uint32_t *group_filter, const uint32_t *group_filter2,
                        const bref_desc *desc, uint64_t *out_words, int n_threads);

/* Q keys x n_units units -> bit matrix, unit-major rows of ceil(Q/64) u64 words:
 * bit q of row u = TestString(key q) on filter desc[u*3+kind[q]] (absent => 1). */
void bref_probe_matrix(const bref_desc *desc, const uint64_t *words, uint64_t n_units,
                       const uint8_t *bytes, const uint64_t *key_off, const uint8_t *kinds,
                       uint32_t n_keys, uint64_t *out_matrix, int n_threads);

/* Candidate mask (bit u of out_mask = unit u survives) via the postfix program
 * whose leaf i is key i.  prog==NULL => every unit survives. */
int bref_probe_mask(const bref_desc *desc, const uint64_t *words, uint64_t n_units,
                    const uint8_t *bytes, const uint64_t *key_off, const uint8_t *kinds,
                    uint32_t n_keys, const bref_op *prog, uint32_t prog_len,
                    uint64_t *out_mask, int n_threads);

/* The Go engine's per-query block loop (query_exec.go:572-615): for every unit,
 * parseFilterSection (CRC32C + BE decode into fresh words) THEN evaluate the
 * expression with short-circuit, hashing each leaf key again per unit.
 * sections = concatenated raw section bytes; unit u = [sec_off[u], sec_off[u+1]).
 * Returns number of sections that failed to parse (those units are kept, like
 * a per-block error).  This is the `--impl reference` / cpu_baseline workload. */
int64_t bref_probe_sections(const uint8_t *sections, const uint64_t *sec_off, uint64_t n_units,
                            const bref_expr *expr, uint64_t *out_mask, int n_threads);

/* Same loop, but for a batch of independent single-condition queries evaluated in
 * one pass over the corpus: per unit, parseFilterSection ONCE, then TestString for
 * every key (hashing the key again for every unit, as TestString does), writing
 * the (unit x key) bit matrix.  This is the conservative CPU statement of the
 * batched probe: the real engine would re-decode every section once per query. */
int64_t bref_probe_sections_matrix(const uint8_t *sections, const uint64_t *sec_off, uint64_t n_units,
                                   const uint8_t *bytes, const uint64_t *key_off, const uint8_t *kinds,
                                   uint32_t n_keys, uint64_t *out_matrix, int n_threads);

/* Encode units (descriptors over a words array) into back-to-back filter sections. */
uint64_t bref_encode_sections(const bref_desc *desc, const uint64_t *words, uint64_t n_units,
                              uint64_t *sec_off, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
