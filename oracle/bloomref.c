/*
 * bloomref.c — CPU ORACLE (test infrastructure only; see bloomref.h header).
 * PARITY STATUS: "parity unpinned" at the bit level (no Go toolchain, library
 * not vendored, no golden bitsets in the reference); murmur3 / CRC32C / (m,k)
 * pinned by public vectors, see tests/test_oracle.py.
 */
#define _GNU_SOURCE
#include "bloomref.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================== *
 * MurmurHash3_x64_128 — bloom/v3 murmur.go restates Austin Appleby's public
 * domain algorithm (constants c1_128/c2_128, bmix, fmix64) with seed 0.
 * ======================================================================== */
#define C1_128 0x87c37b91114253d5ULL
#define C2_128 0x4cf5ad432745937fULL

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

static inline uint64_t load_le64(const uint8_t *p) {
    return (uint64_t)p[0] | ((uint64_t)p[1] << 8) | ((uint64_t)p[2] << 16) |
           ((uint64_t)p[3] << 24) | ((uint64_t)p[4] << 32) | ((uint64_t)p[5] << 40) |
           ((uint64_t)p[6] << 48) | ((uint64_t)p[7] << 56);
}

/* murmur.go bmix_words */
static inline void bmix_words(uint64_t *h1, uint64_t *h2, uint64_t k1, uint64_t k2) {
    k1 *= C1_128; k1 = rotl64(k1, 31); k1 *= C2_128; *h1 ^= k1;
    *h1 = rotl64(*h1, 27); *h1 += *h2; *h1 = *h1 * 5 + 0x52dce729;
    k2 *= C2_128; k2 = rotl64(k2, 33); k2 *= C1_128; *h2 ^= k2;
    *h2 = rotl64(*h2, 31); *h2 += *h1; *h2 = *h2 * 5 + 0x38495ab5;
}

/* murmur.go sum128: tail (< 16 bytes) + finalisation */
static inline void tail_and_final(uint64_t h1, uint64_t h2, const uint8_t *tail, size_t tail_len,
                                  uint64_t total_len, uint64_t out[2]) {
    uint64_t k1 = 0, k2 = 0;
    for (size_t i = tail_len; i > 8; i--) k2 ^= (uint64_t)tail[i - 1] << (8 * (i - 9));
    if (tail_len > 8) { k2 *= C2_128; k2 = rotl64(k2, 33); k2 *= C1_128; h2 ^= k2; }
    size_t n1 = tail_len > 8 ? 8 : tail_len;
    for (size_t i = n1; i > 0; i--) k1 ^= (uint64_t)tail[i - 1] << (8 * (i - 1));
    if (tail_len > 0) { k1 *= C1_128; k1 = rotl64(k1, 31); k1 *= C2_128; h1 ^= k1; }
    h1 ^= total_len; h2 ^= total_len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

void bref_murmur3_x64_128(const void *data, size_t len, uint32_t seed, uint64_t out[2]) {
    const uint8_t *p = (const uint8_t *)data;
    uint64_t h1 = seed, h2 = seed;
    size_t nblocks = len / 16;
    for (size_t i = 0; i < nblocks; i++) bmix_words(&h1, &h2, load_le64(p + 16 * i), load_le64(p + 16 * i + 8));
    tail_and_final(h1, h2, p + 16 * nblocks, len & 15, (uint64_t)len, out);
}

/* bloom.go baseHashes -> murmur.go sum256: "strictly equivalent to"
 *   hasher.Write(data); v1,v2 := Sum128(); hasher.Write([]byte{1}); v3,v4 := Sum128()
 * i.e. murmur(data) and murmur(data || 0x01), both seed 0, sharing the block phase. */
void bref_base_hashes(const uint8_t *data, size_t len, uint64_t h[4]) {
    uint64_t h1 = 0, h2 = 0;
    size_t nblocks = len / 16;
    for (size_t i = 0; i < nblocks; i++)
        bmix_words(&h1, &h2, load_le64(data + 16 * i), load_le64(data + 16 * i + 8));
    size_t tail_len = len & 15;
    const uint8_t *tail = data + 16 * nblocks;
    tail_and_final(h1, h2, tail, tail_len, (uint64_t)len, h);
    uint8_t ext[16];
    memcpy(ext, tail, tail_len);
    ext[tail_len] = 1;
    if (tail_len + 1 == 16) { /* the virtual byte completes a block: empty tail */
        bmix_words(&h1, &h2, load_le64(ext), load_le64(ext + 8));
        tail_and_final(h1, h2, ext, 0, (uint64_t)len + 1, h + 2);
    } else {
        tail_and_final(h1, h2, ext, tail_len + 1, (uint64_t)len + 1, h + 2);
    }
}

uint64_t bref_location(const uint64_t h[4], uint64_t i) {
    return h[i % 2] + i * h[2 + (((i + (i % 2)) % 4) / 2)];
}

void bref_estimate_parameters(uint64_t n, double p, uint64_t *m, uint64_t *k) {
    /* bloom.go EstimateParameters:
     *   m = uint(math.Ceil(-1 * float64(n) * math.Log(p) / math.Pow(math.Log(2), 2)))
     *   k = uint(math.Ceil(math.Log(2) * float64(m) / float64(n)))
     * then New() clamps both to >= 1. */
    double ln2 = log(2.0);
    double md = ceil(-1.0 * (double)n * log(p) / pow(ln2, 2.0));
    uint64_t mm = (uint64_t)md;
    double kd = ceil(ln2 * (double)mm / (double)n);
    uint64_t kk = (uint64_t)kd;
    *m = mm < 1 ? 1 : mm;
    *k = kk < 1 ? 1 : kk;
}

/* ======================================================================== *
 * BloomFilter / BitSet
 * ======================================================================== */
bref_filter *bref_filter_new(uint64_t m, uint64_t k) {
    bref_filter *f = (bref_filter *)calloc(1, sizeof(*f));
    if (!f) return NULL;
    f->m = m < 1 ? 1 : m;
    f->k = k < 1 ? 1 : k;
    f->nwords = (f->m + 63) >> 6;
    f->words = (uint64_t *)calloc(f->nwords ? f->nwords : 1, sizeof(uint64_t));
    if (!f->words) { free(f); return NULL; }
    return f;
}

bref_filter *bref_filter_new_with_estimates(uint64_t n, double fpr) {
    uint64_t m, k;
    bref_estimate_parameters(n, fpr, &m, &k);
    return bref_filter_new(m, k);
}

void bref_filter_free(bref_filter *f) {
    if (!f) return;
    free(f->words);
    free(f);
}

void bref_filter_add(bref_filter *f, const uint8_t *data, size_t len) {
    uint64_t h[4];
    bref_base_hashes(data, len, h);
    for (uint64_t i = 0; i < f->k; i++) {
        uint64_t bit = bref_location(h, i) % f->m;
        f->words[bit >> 6] |= 1ULL << (bit & 63);
    }
}

static inline int test_hashes(const uint64_t m, const uint64_t k, const uint64_t *words, const uint64_t h[4]) {
    for (uint64_t i = 0; i < k; i++) {
        uint64_t bit = bref_location(h, i) % m;
        if (!(words[bit >> 6] & (1ULL << (bit & 63)))) return 0;
    }
    return 1;
}

int bref_filter_test(const bref_filter *f, const uint8_t *data, size_t len) {
    uint64_t h[4];
    bref_base_hashes(data, len, h);
    return test_hashes(f->m, f->k, f->words, h);
}

int bref_filter_equal(const bref_filter *a, const bref_filter *b) {
    if (!a || !b) return a == b;
    return a->m == b->m && a->k == b->k && a->nwords == b->nwords &&
           memcmp(a->words, b->words, a->nwords * 8) == 0;
}

uint64_t bref_filter_m(const bref_filter *f) { return f->m; }
uint64_t bref_filter_k(const bref_filter *f) { return f->k; }
uint64_t bref_filter_nwords(const bref_filter *f) { return f->nwords; }
uint64_t *bref_filter_words(bref_filter *f) { return f->words; }

static inline void put_be64(uint8_t *p, uint64_t v) {
    for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (56 - 8 * i));
}
static inline uint64_t get_be64(const uint8_t *p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return __builtin_bswap64(v); /* binary.BigEndian.Uint64 compiles to MOVBE/BSWAP in Go too */
}

size_t bref_filter_serialized_size(const bref_filter *f) { return 24 + 8 * f->nwords; }

size_t bref_filter_write_to(const bref_filter *f, uint8_t *out) {
    put_be64(out, f->m);
    put_be64(out + 8, f->k);
    put_be64(out + 16, f->m); /* bitset length */
    for (uint64_t i = 0; i < f->nwords; i++) put_be64(out + 24 + 8 * i, f->words[i]);
    return 24 + 8 * f->nwords;
}

bref_filter *bref_filter_read_from(const uint8_t *in, size_t len, size_t *consumed) {
    if (len < 24) return NULL;
    uint64_t m = get_be64(in), k = get_be64(in + 8), bitlen = get_be64(in + 16);
    uint64_t nwords = (bitlen + 63) >> 6;
    if (bitlen > (1ULL << 40) || len < 24 + 8 * nwords) return NULL;
    bref_filter *f = (bref_filter *)calloc(1, sizeof(*f));
    if (!f) return NULL;
    f->m = m; f->k = k; f->nwords = nwords;
    f->words = (uint64_t *)malloc((nwords ? nwords : 1) * 8);
    if (!f->words) { free(f); return NULL; }
    for (uint64_t i = 0; i < nwords; i++) f->words[i] = get_be64(in + 24 + 8 * i);
    if (consumed) *consumed = 24 + 8 * nwords;
    return f;
}

bref_filter *bref_build_sized_filter(const uint8_t *bytes, const uint64_t *key_off,
                                     uint64_t n_keys, double fpr) {
    /* ingest.go:139-145: NewWithEstimates(max(len(entries),1), fpr); AddString each */
    bref_filter *f = bref_filter_new_with_estimates(n_keys > 1 ? n_keys : 1, fpr);
    if (!f) return NULL;
    for (uint64_t i = 0; i < n_keys; i++)
        bref_filter_add(f, bytes + key_off[i], (size_t)(key_off[i + 1] - key_off[i]));
    return f;
}

/* ======================================================================== *
 * CRC32C (Castagnoli, reflected poly 0x82F63B78) — file_format.go:44 crc32cTable
 * ======================================================================== */
static uint32_t crc32c_table[8][256];
static pthread_once_t crc_once = PTHREAD_ONCE_INIT;
static void crc32c_init(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int j = 0; j < 8; j++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
        crc32c_table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++)
            crc32c_table[t][i] = (crc32c_table[t - 1][i] >> 8) ^ crc32c_table[0][crc32c_table[t - 1][i] & 0xff];
}

#if defined(__x86_64__)
/* Go's hash/crc32 uses the SSE4.2 crc32 instruction for the Castagnoli table; so does the CPU
 * baseline when the host has it (same polynomial, same result as the table path below). */
__attribute__((target("sse4.2")))
static uint32_t crc32c_hw(const uint8_t *data, size_t len) {
    uint64_t c = 0xFFFFFFFFu;
    while (len >= 8) {
        uint64_t v;
        memcpy(&v, data, 8);
        c = __builtin_ia32_crc32di(c, v);
        data += 8; len -= 8;
    }
    uint32_t c32 = (uint32_t)c;
    while (len--) c32 = __builtin_ia32_crc32qi(c32, *data++);
    return c32 ^ 0xFFFFFFFFu;
}
#endif

uint32_t bref_crc32c(const uint8_t *data, size_t len) {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("sse4.2")) return crc32c_hw(data, len);
#endif
    return bref_crc32c_sw(data, len);
}

/* The restated algorithm proper (table-driven); bref_crc32c falls back to it where the instruction is missing,
 * and tests/test_oracle.py checks it against the instruction (oracle/pins/crc32c_hw.c) where it is present. */
uint32_t bref_crc32c_sw(const uint8_t *data, size_t len) {
    pthread_once(&crc_once, crc32c_init);
    uint32_t c = 0xFFFFFFFFu;
    while (len >= 8) { /* slicing-by-8, comparable to Go's software path */
        uint32_t lo = ((uint32_t)data[0] | (uint32_t)data[1] << 8 | (uint32_t)data[2] << 16 | (uint32_t)data[3] << 24) ^ c;
        uint32_t hi = (uint32_t)data[4] | (uint32_t)data[5] << 8 | (uint32_t)data[6] << 16 | (uint32_t)data[7] << 24;
        c = crc32c_table[7][lo & 0xff] ^ crc32c_table[6][(lo >> 8) & 0xff] ^
            crc32c_table[5][(lo >> 16) & 0xff] ^ crc32c_table[4][lo >> 24] ^
            crc32c_table[3][hi & 0xff] ^ crc32c_table[2][(hi >> 8) & 0xff] ^
            crc32c_table[1][(hi >> 16) & 0xff] ^ crc32c_table[0][hi >> 24];
        data += 8; len -= 8;
    }
    while (len--) c = crc32c_table[0][(c ^ *data++) & 0xff] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

/* ======================================================================== *
 * Filter section codec — file_format.go:343-385 / 392-448
 * ======================================================================== */
static inline void put_le32(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
static inline uint32_t get_le32(const uint8_t *p) {
    return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}

size_t bref_section_encode(const bref_filter *const filters[3], uint8_t *out) {
    size_t size = 1;
    uint8_t flags = 0;
    for (int i = 0; i < 3; i++)
        if (filters && filters[i]) { flags |= (uint8_t)(1u << i); size += 4 + bref_filter_serialized_size(filters[i]); }
    size += 4;
    if (!out) return size;
    uint8_t *p = out;
    *p++ = flags;
    for (int i = 0; i < 3; i++) {
        if (!(flags & (1u << i))) continue;
        size_t n = bref_filter_serialized_size(filters[i]);
        put_le32(p, (uint32_t)n); p += 4;
        bref_filter_write_to(filters[i], p); p += n;
    }
    put_le32(p, bref_crc32c(out, (size_t)(p - out)));
    return size;
}

int bref_section_parse(const uint8_t *section, size_t len, bref_filter *filters[3]) {
    filters[0] = filters[1] = filters[2] = NULL;
    if (len < 4 + 1) return -1;
    size_t plen = len - 4;
    if (bref_crc32c(section, plen) != get_le32(section + plen)) return -2;
    uint8_t flags = section[0];
    if (flags & ~7u) return -3;
    const uint8_t *rest = section + 1;
    size_t remaining = plen - 1;
    int rc = 0;
    for (int i = 0; i < 3 && rc == 0; i++) {
        if (!(flags & (1u << i))) continue;
        if (remaining < 4) { rc = -4; break; }
        uint32_t n = get_le32(rest); rest += 4; remaining -= 4;
        if ((uint64_t)n > (uint64_t)remaining) { rc = -5; break; }
        size_t used = 0;
        filters[i] = bref_filter_read_from(rest, n, &used);
        if (!filters[i]) { rc = -6; break; }
        rest += n; remaining -= n;
    }
    if (rc == 0 && remaining != 0) rc = -7;
    if (rc != 0)
        for (int i = 0; i < 3; i++) { bref_filter_free(filters[i]); filters[i] = NULL; }
    return rc;
}

/* ======================================================================== *
 * Expression tree — query_exec.go:75-159
 * ======================================================================== */
static int eval_condition(const bref_filter *ff, const bref_filter *tf, const bref_filter *ftf,
                          const bref_expr *e) {
    switch (e->cond_type) {
    case BREF_COND_FIELD:
        if (!ff) return 1; /* nil filter cannot disqualify (query_exec.go:137-140) */
        return bref_filter_test(ff, e->field, (size_t)e->field_len);
    case BREF_COND_TOKEN:
        if (!tf) return 1;
        return bref_filter_test(tf, e->token, (size_t)e->token_len);
    case BREF_COND_FIELD_TOKEN: {
        if (!ftf) return 1;
        /* makeFieldTokenKey (tokenizer.go:509): field + "::" + token */
        size_t n = (size_t)(e->field_len + 2 + e->token_len);
        uint8_t stackbuf[256];
        uint8_t *key = n <= sizeof(stackbuf) ? stackbuf : (uint8_t *)malloc(n);
        memcpy(key, e->field, (size_t)e->field_len);
        key[e->field_len] = ':'; key[e->field_len + 1] = ':';
        memcpy(key + e->field_len + 2, e->token, (size_t)e->token_len);
        int r = bref_filter_test(ftf, key, n);
        if (key != stackbuf) free(key);
        return r;
    }
    default:
        return 0; /* unknown condition type (query_exec.go:155-156) */
    }
}

static int eval_expr(const bref_filter *ff, const bref_filter *tf, const bref_filter *ftf,
                     const bref_expr *e) {
    if (!e) return 1;
    switch (e->type) {
    case BREF_EXPR_CONDITION:
        if (!e->has_condition) return 1;
        return eval_condition(ff, tf, ftf, e);
    case BREF_EXPR_OR:
        if (e->n_children == 0) return 0;
        for (int i = 0; i < e->n_children; i++)
            if (eval_expr(ff, tf, ftf, &e->children[i])) return 1;
        return 0;
    case BREF_EXPR_AND:
        for (int i = 0; i < e->n_children; i++)
            if (!eval_expr(ff, tf, ftf, &e->children[i])) return 0;
        return 1;
    default:
        return 0;
    }
}

int bref_evaluate_bloom_filters(const bref_filter *field_f, const bref_filter *token_f,
                                const bref_filter *fieldtoken_f, const bref_expr *expr) {
    if (!expr) return 1;
    return eval_expr(field_f, token_f, fieldtoken_f, expr);
}

int bref_eval_postfix(const bref_op *prog, uint32_t prog_len, const uint8_t *leaf_bits,
                      uint32_t n_leaves) {
    uint8_t stack[256];
    uint32_t sp = 0;
    for (uint32_t pc = 0; pc < prog_len; pc++) {
        uint32_t arg = prog[pc].arg;
        switch (prog[pc].op) {
        case BREF_OP_LEAF:
            if (arg >= n_leaves || sp >= sizeof(stack)) return -1;
            stack[sp++] = leaf_bits[arg] ? 1 : 0;
            break;
        case BREF_OP_TRUE: case BREF_OP_FALSE:
            if (sp >= sizeof(stack)) return -1;
            stack[sp++] = prog[pc].op == BREF_OP_TRUE;
            break;
        case BREF_OP_AND: case BREF_OP_OR: {
            if (arg > sp) return -1;
            uint8_t v = prog[pc].op == BREF_OP_AND; /* empty AND true, empty OR false */
            for (uint32_t i = 0; i < arg; i++) {
                uint8_t c = stack[--sp];
                v = prog[pc].op == BREF_OP_AND ? (v & c) : (v | c);
            }
            if (sp >= sizeof(stack)) return -1;
            stack[sp++] = v;
            break;
        }
        default:
            return -1;
        }
    }
    return sp == 1 ? stack[0] : -1;
}

/* ======================================================================== *
 * Bulk helpers (pthread fan-out over groups / units)
 * ======================================================================== */
typedef void (*range_fn)(void *ctx, uint64_t lo, uint64_t hi);
typedef struct { range_fn fn; void *ctx; uint64_t lo, hi; } range_task;
static void *range_thread(void *p) {
    range_task *t = (range_task *)p;
    t->fn(t->ctx, t->lo, t->hi);
    return NULL;
}
static void parallel_ranges(uint64_t n, int n_threads, range_fn fn, void *ctx) {
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n) n_threads = n ? (int)n : 1;
    if (n_threads == 1) { fn(ctx, 0, n); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    range_task *tk = (range_task *)malloc(sizeof(range_task) * (size_t)n_threads);
    for (int i = 0; i < n_threads; i++) {
        tk[i].fn = fn; tk[i].ctx = ctx;
        tk[i].lo = n * (uint64_t)i / (uint64_t)n_threads;
        tk[i].hi = n * (uint64_t)(i + 1) / (uint64_t)n_threads;
        pthread_create(&th[i], NULL, range_thread, &tk[i]);
    }
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th); free(tk);
}

typedef struct {
    const uint8_t *bytes; const uint64_t *key_off; const uint64_t *group_begin;
    const uint32_t *group_filter, *group_filter2; const bref_desc *desc; uint64_t *out;
    int atomic2;
} build_ctx;

static void build_range(void *p, uint64_t lo, uint64_t hi) {
    build_ctx *c = (build_ctx *)p;
    for (uint64_t g = lo; g < hi; g++) {
        const bref_desc *d1 = &c->desc[c->group_filter[g]];
        const bref_desc *d2 = NULL;
        if (c->group_filter2 && c->group_filter2[g] != 0xFFFFFFFFu) d2 = &c->desc[c->group_filter2[g]];
        for (uint64_t i = c->group_begin[g]; i < c->group_begin[g + 1]; i++) {
            uint64_t h[4];
            bref_base_hashes(c->bytes + c->key_off[i], (size_t)(c->key_off[i + 1] - c->key_off[i]), h);
            for (uint64_t j = 0; j < d1->k; j++) {
                uint64_t bit = bref_location(h, j) % d1->m;
                c->out[d1->word_off + (bit >> 6)] |= 1ULL << (bit & 63);
            }
            if (d2)
                for (uint64_t j = 0; j < d2->k; j++) {
                    uint64_t bit = bref_location(h, j) % d2->m;
                    uint64_t mask = 1ULL << (bit & 63);
                    if (c->atomic2) __atomic_fetch_or(&c->out[d2->word_off + (bit >> 6)], mask, __ATOMIC_RELAXED);
                    else c->out[d2->word_off + (bit >> 6)] |= mask;
                }
        }
    }
}

void bref_build_filters(const uint8_t *bytes, const uint64_t *key_off,
                        const uint64_t *group_begin, uint32_t n_groups,
                        const uint32_t *group_filter, const uint32_t *group_filter2,
                        const bref_desc *desc, uint64_t *out_words, int n_threads) {
    /* NOTE: with n_threads > 1 every primary filter must be owned by one group
     * (true for block filters); shared secondary (file-level) filters use atomics. */
    build_ctx c = { bytes, key_off, group_begin, group_filter, group_filter2, desc, out_words, n_threads > 1 };
    parallel_ranges(n_groups, n_threads, build_range, &c);
}

typedef struct {
    const bref_desc *desc; const uint64_t *words; const uint64_t (*hashes)[4];
    const uint8_t *kinds; uint32_t n_keys; uint64_t *out; uint64_t row_words;
    const bref_op *prog; uint32_t prog_len; uint64_t *mask; int bad;
} probe_ctx;

static void probe_range(void *p, uint64_t lo, uint64_t hi) {
    probe_ctx *c = (probe_ctx *)p;
    uint8_t *bits = (uint8_t *)malloc(c->n_keys ? c->n_keys : 1);
    for (uint64_t u = lo; u < hi; u++) {
        for (uint32_t q = 0; q < c->n_keys; q++) {
            const bref_desc *d = &c->desc[u * 3 + c->kinds[q]];
            bits[q] = d->m == 0 ? 1 : (uint8_t)test_hashes(d->m, d->k, c->words + d->word_off, c->hashes[q]);
        }
        if (c->out)
            for (uint32_t q = 0; q < c->n_keys; q++)
                if (bits[q]) c->out[u * c->row_words + (q >> 6)] |= 1ULL << (q & 63);
        if (c->mask) {
            int v = c->prog ? bref_eval_postfix(c->prog, c->prog_len, bits, c->n_keys) : 1;
            if (v < 0) { c->bad = 1; v = 0; }
            if (v) __atomic_fetch_or(&c->mask[u >> 6], 1ULL << (u & 63), __ATOMIC_RELAXED);
        }
    }
    free(bits);
}

static uint64_t (*hash_keys(const uint8_t *bytes, const uint64_t *key_off, uint32_t n))[4] {
    uint64_t (*h)[4] = (uint64_t (*)[4])malloc(sizeof(uint64_t[4]) * (n ? n : 1));
    for (uint32_t q = 0; q < n; q++)
        bref_base_hashes(bytes + key_off[q], (size_t)(key_off[q + 1] - key_off[q]), h[q]);
    return h;
}

void bref_probe_matrix(const bref_desc *desc, const uint64_t *words, uint64_t n_units,
                       const uint8_t *bytes, const uint64_t *key_off, const uint8_t *kinds,
                       uint32_t n_keys, uint64_t *out_matrix, int n_threads) {
    uint64_t row_words = ((uint64_t)n_keys + 63) / 64;
    memset(out_matrix, 0, (size_t)(n_units * row_words * 8));
    uint64_t (*h)[4] = hash_keys(bytes, key_off, n_keys);
    probe_ctx c = { desc, words, (const uint64_t (*)[4])h, kinds, n_keys, out_matrix, row_words, NULL, 0, NULL, 0 };
    parallel_ranges(n_units, n_threads, probe_range, &c);
    free(h);
}

int bref_probe_mask(const bref_desc *desc, const uint64_t *words, uint64_t n_units,
                    const uint8_t *bytes, const uint64_t *key_off, const uint8_t *kinds,
                    uint32_t n_keys, const bref_op *prog, uint32_t prog_len,
                    uint64_t *out_mask, int n_threads) {
    memset(out_mask, 0, (size_t)(((n_units + 63) / 64) * 8));
    uint64_t (*h)[4] = hash_keys(bytes, key_off, n_keys);
    probe_ctx c = { desc, words, (const uint64_t (*)[4])h, kinds, n_keys, NULL, 0, prog, prog_len, out_mask, 0 };
    parallel_ranges(n_units, n_threads, probe_range, &c);
    free(h);
    return c.bad ? -1 : 0;
}

typedef struct {
    const uint8_t *sections; const uint64_t *sec_off; const bref_expr *expr;
    uint64_t *mask; int64_t errors;
} sec_ctx;

static void sections_range(void *p, uint64_t lo, uint64_t hi) {
    sec_ctx *c = (sec_ctx *)p;
    int64_t errs = 0;
    for (uint64_t u = lo; u < hi; u++) {
        bref_filter *f[3];
        int rc = bref_section_parse(c->sections + c->sec_off[u], (size_t)(c->sec_off[u + 1] - c->sec_off[u]), f);
        /* query_exec.go:580-590: a section that fails to parse is an ERROR for that block — it is recorded
         * and the loop `continue`s: the block is not a candidate (never scanned), not a survivor. */
        int keep = 0;
        if (rc != 0) errs++;
        else {
            keep = bref_evaluate_bloom_filters(f[0], f[1], f[2], c->expr);
            for (int i = 0; i < 3; i++) bref_filter_free(f[i]);
        }
        if (keep) __atomic_fetch_or(&c->mask[u >> 6], 1ULL << (u & 63), __ATOMIC_RELAXED);
    }
    __atomic_fetch_add(&c->errors, errs, __ATOMIC_RELAXED);
}

int64_t bref_probe_sections(const uint8_t *sections, const uint64_t *sec_off, uint64_t n_units,
                            const bref_expr *expr, uint64_t *out_mask, int n_threads) {
    memset(out_mask, 0, (size_t)(((n_units + 63) / 64) * 8));
    sec_ctx c = { sections, sec_off, expr, out_mask, 0 };
    parallel_ranges(n_units, n_threads, sections_range, &c);
    return c.errors;
}

typedef struct {
    const uint8_t *sections; const uint64_t *sec_off;
    const uint8_t *bytes; const uint64_t *key_off; const uint8_t *kinds; uint32_t n_keys;
    uint64_t *out; uint64_t row_words; int64_t errors;
} secm_ctx;

static void sections_matrix_range(void *p, uint64_t lo, uint64_t hi) {
    secm_ctx *c = (secm_ctx *)p;
    int64_t errs = 0;
    for (uint64_t u = lo; u < hi; u++) {
        bref_filter *f[3];
        int rc = bref_section_parse(c->sections + c->sec_off[u], (size_t)(c->sec_off[u + 1] - c->sec_off[u]), f);
        if (rc != 0) { errs++; f[0] = f[1] = f[2] = NULL; }
        uint64_t *row = c->out + u * c->row_words;
        for (uint32_t q = 0; q < c->n_keys; q++) {
            const bref_filter *flt = f[c->kinds[q]];
            int r = flt ? bref_filter_test(flt, c->bytes + c->key_off[q], (size_t)(c->key_off[q + 1] - c->key_off[q])) : 1;
            if (r) row[q >> 6] |= 1ULL << (q & 63);
        }
        for (int i = 0; i < 3; i++) bref_filter_free(f[i]);
    }
    __atomic_fetch_add(&c->errors, errs, __ATOMIC_RELAXED);
}

int64_t bref_probe_sections_matrix(const uint8_t *sections, const uint64_t *sec_off, uint64_t n_units,
                                   const uint8_t *bytes, const uint64_t *key_off, const uint8_t *kinds,
                                   uint32_t n_keys, uint64_t *out_matrix, int n_threads) {
    uint64_t row_words = ((uint64_t)n_keys + 63) / 64;
    memset(out_matrix, 0, (size_t)(n_units * row_words * 8));
    secm_ctx c = { sections, sec_off, bytes, key_off, kinds, n_keys, out_matrix, row_words, 0 };
    parallel_ranges(n_units, n_threads, sections_matrix_range, &c);
    return c.errors;
}

/* Encode every unit's (up to three) filters, given as descriptors over a words array,
 * into back-to-back filter sections (file_format.go:343-385).  Two-pass: call with
 * out == NULL to obtain sec_off (n_units+1) and the total size. */
uint64_t bref_encode_sections(const bref_desc *desc, const uint64_t *words, uint64_t n_units,
                              uint64_t *sec_off, uint8_t *out) {
    uint64_t pos = 0;
    for (uint64_t u = 0; u < n_units; u++) {
        bref_filter tmp[3];
        const bref_filter *fl[3] = {NULL, NULL, NULL};
        for (int k = 0; k < 3; k++) {
            const bref_desc *d = &desc[u * 3 + k];
            if (d->m == 0) continue;
            tmp[k].m = d->m; tmp[k].k = d->k; tmp[k].nwords = (d->m + 63) >> 6;
            tmp[k].words = (uint64_t *)(words + d->word_off);
            fl[k] = &tmp[k];
        }
        sec_off[u] = pos;
        pos += bref_section_encode(fl, out ? out + pos : NULL);
    }
    sec_off[n_units] = pos;
    return pos;
}
