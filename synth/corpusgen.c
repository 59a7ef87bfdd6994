/*
 * corpusgen.c — synthetic log-corpus entry sets: INPUT DATA for tests and bench
 * (not part of the oracle — it computes no bloom arithmetic — and never linked into the product).
 *
 * Restates, for the one fixed row shape of the reference's benchmarks
 * (benchRows, bench_test.go:16-49), what bloomEntrySets.indexRow collects
 * (ingest.go:55-102 via the path walker row_matcher.go:56-99 and the
 * whitespace/lower tokenizer tokenizer.go:141-143):
 *   fields      : every path incl. intermediate object paths  -> 9 keys
 *   tokens      : whitespace-split lower-cased leaf text; numbers by raw literal
 *   fieldTokens : exact-leaf-path "::" token  (arrays contribute at the array's path)
 * Row i: timestamp 1700000000+i, level in 4, service in 5, message = 8 of 13 words,
 * user_id in [0,100000), nested.region "region-%d" in 8, nested.az "az-%d" in 3,
 * tags = 2 of 13 words.  Go's math/rand stream (seed 42) cannot be reproduced
 * without Go, so draws come from a documented SplitMix64 stream per block
 * (state = seed + (block+1)*0x9E3779B97F4A7C15), value % n.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const char *kLevels[4] = {"debug", "info", "warn", "error"};
static const char *kServices[5] = {"auth", "payment", "search", "gateway", "billing"};
static const char *kWords[13] = {"connection", "timeout", "retry", "database", "request", "processed", "failed",
                                 "succeeded", "cache", "miss", "upstream", "latency", "shard"};
static const char *kFields[9] = {"timestamp", "level", "service", "message", "user_id",
                                 "nested", "nested.region", "nested.az", "tags"};

static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

typedef struct {
    uint8_t *bytes; uint64_t nbytes, cap_bytes;
    uint64_t *off; uint64_t nkeys, cap_keys;
} keybuf;

static void kb_reserve(keybuf *b, uint64_t add_bytes, uint64_t add_keys) {
    if (b->nbytes + add_bytes > b->cap_bytes) {
        while (b->nbytes + add_bytes > b->cap_bytes) b->cap_bytes = b->cap_bytes ? b->cap_bytes * 2 : 1 << 20;
        b->bytes = (uint8_t *)realloc(b->bytes, b->cap_bytes);
    }
    if (b->nkeys + add_keys + 1 > b->cap_keys) {
        while (b->nkeys + add_keys + 1 > b->cap_keys) b->cap_keys = b->cap_keys ? b->cap_keys * 2 : 1 << 16;
        b->off = (uint64_t *)realloc(b->off, b->cap_keys * sizeof(uint64_t));
    }
}
static void kb_add2(keybuf *b, const char *a, size_t la, const char *c, size_t lc) {
    kb_reserve(b, la + lc, 1);
    if (b->nkeys == 0) b->off[0] = 0;
    memcpy(b->bytes + b->nbytes, a, la);
    memcpy(b->bytes + b->nbytes + la, c, lc);
    b->nbytes += la + lc;
    b->off[++b->nkeys] = b->nbytes;
}
static inline void u64_to_dec(uint64_t v, char *out) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < n; i++) out[i] = tmp[n - 1 - i];
    out[n] = 0;
}
static void kb_add(keybuf *b, const char *a) { kb_add2(b, a, strlen(a), "", 0); }
static void kb_addft(keybuf *b, const char *path, const char *tok) {
    char tmp[96];
    size_t n = strlen(path);
    memcpy(tmp, path, n);
    tmp[n] = ':'; tmp[n + 1] = ':';
    kb_add2(b, tmp, n + 2, tok, strlen(tok));
}

typedef struct bgen_corpus {
    uint8_t *bytes;          /* packed keys */
    uint64_t *key_off;       /* n_keys + 1 */
    uint64_t n_keys;
    uint64_t *group_begin;   /* 3*n_blocks + 1; group 3*b+kind */
    uint64_t n_blocks;
    uint64_t *file_counts;   /* n_files * 3 exact distinct counts of each file's union */
    uint64_t n_files;
} bgen_corpus;

void bgen_free(bgen_corpus *c) {
    if (!c) return;
    free(c->bytes); free(c->key_off); free(c->group_begin); free(c->file_counts); free(c);
}

/* Blocks [block_lo, block_lo+n_blocks), rows_per_block rows each; files are
 * blocks_per_file consecutive blocks (n_blocks must be a multiple). */
bgen_corpus *bgen_generate(uint64_t seed, uint64_t block_lo, uint64_t n_blocks, uint32_t rows_per_block,
                           uint32_t blocks_per_file) {
    if (blocks_per_file == 0 || n_blocks % blocks_per_file) return NULL;
    bgen_corpus *c = (bgen_corpus *)calloc(1, sizeof(*c));
    keybuf kb = {0};
    c->n_blocks = n_blocks;
    c->n_files = n_blocks / blocks_per_file;
    c->group_begin = (uint64_t *)malloc((3 * n_blocks + 1) * sizeof(uint64_t));
    size_t n_fc = (size_t)c->n_files * 3;
    if (n_fc == 0) n_fc = 1;
    c->file_counts = (uint64_t *)calloc(n_fc, sizeof(uint64_t));
    uint8_t *uid_seen = (uint8_t *)malloc(100000), *uid_file = (uint8_t *)malloc(100000);
    uint32_t *uids = (uint32_t *)malloc(sizeof(uint32_t) * (rows_per_block ? rows_per_block : 1));
    char num[32], tok[32];
    kb_reserve(&kb, 0, 0);
    kb.off[0] = 0;
    uint8_t f_lvl[4], f_svc[5], f_msg[13], f_tag[13], f_reg[8], f_az[3];
    uint64_t file_uid_count = 0;
    for (uint64_t bi = 0; bi < n_blocks; bi++) {
        const uint64_t b = block_lo + bi;
        if (bi % blocks_per_file == 0) {
            memset(uid_file, 0, 100000);
            memset(f_lvl, 0, 4); memset(f_svc, 0, 5); memset(f_msg, 0, 13); memset(f_tag, 0, 13);
            memset(f_reg, 0, 8); memset(f_az, 0, 3);
            file_uid_count = 0;
        }
        uint64_t st = seed + (b + 1) * 0x9E3779B97F4A7C15ULL;
        uint8_t lvl[4] = {0}, svc[5] = {0}, wmsg[13] = {0}, wtag[13] = {0}, reg[8] = {0}, az[3] = {0};
        memset(uid_seen, 0, 100000);
        uint32_t n_uid = 0;
        for (uint32_t j = 0; j < rows_per_block; j++) {
            lvl[splitmix64(&st) % 4] = 1;
            svc[splitmix64(&st) % 5] = 1;
            for (int w = 0; w < 8; w++) wmsg[splitmix64(&st) % 13] = 1;
            uint32_t uid = (uint32_t)(splitmix64(&st) % 100000);
            if (!uid_seen[uid]) { uid_seen[uid] = 1; uids[n_uid++] = uid; }
            reg[splitmix64(&st) % 8] = 1;
            az[splitmix64(&st) % 3] = 1;
            wtag[splitmix64(&st) % 13] = 1;
            wtag[splitmix64(&st) % 13] = 1;
        }
        const uint64_t row0 = b * (uint64_t)rows_per_block;
        /* ---- fields ---- */
        c->group_begin[3 * bi + 0] = kb.nkeys;
        if (rows_per_block) for (int i = 0; i < 9; i++) kb_add(&kb, kFields[i]);
        /* ---- tokens ---- */
        c->group_begin[3 * bi + 1] = kb.nkeys;
        for (uint32_t j = 0; j < rows_per_block; j++) {
            u64_to_dec(1700000000ULL + row0 + j, num);
            kb_add(&kb, num);
        }
        for (uint32_t i = 0; i < n_uid; i++) { u64_to_dec(uids[i], num); kb_add(&kb, num); }
        for (int i = 0; i < 4; i++) if (lvl[i]) kb_add(&kb, kLevels[i]);
        for (int i = 0; i < 5; i++) if (svc[i]) kb_add(&kb, kServices[i]);
        for (int i = 0; i < 13; i++) if (wmsg[i] || wtag[i]) kb_add(&kb, kWords[i]);
        for (int i = 0; i < 8; i++) if (reg[i]) { snprintf(tok, sizeof(tok), "region-%d", i); kb_add(&kb, tok); }
        for (int i = 0; i < 3; i++) if (az[i]) { snprintf(tok, sizeof(tok), "az-%d", i); kb_add(&kb, tok); }
        /* ---- field::token ---- */
        c->group_begin[3 * bi + 2] = kb.nkeys;
        for (uint32_t j = 0; j < rows_per_block; j++) {
            u64_to_dec(1700000000ULL + row0 + j, num);
            kb_addft(&kb, "timestamp", num);
        }
        for (uint32_t i = 0; i < n_uid; i++) { u64_to_dec(uids[i], num); kb_addft(&kb, "user_id", num); }
        for (int i = 0; i < 4; i++) if (lvl[i]) kb_addft(&kb, "level", kLevels[i]);
        for (int i = 0; i < 5; i++) if (svc[i]) kb_addft(&kb, "service", kServices[i]);
        for (int i = 0; i < 13; i++) if (wmsg[i]) kb_addft(&kb, "message", kWords[i]);
        for (int i = 0; i < 13; i++) if (wtag[i]) kb_addft(&kb, "tags", kWords[i]);
        for (int i = 0; i < 8; i++) if (reg[i]) { snprintf(tok, sizeof(tok), "region-%d", i); kb_addft(&kb, "nested.region", tok); }
        for (int i = 0; i < 3; i++) if (az[i]) { snprintf(tok, sizeof(tok), "az-%d", i); kb_addft(&kb, "nested.az", tok); }
        /* ---- file-level union bookkeeping (what unionInto + counts() give, flush.go:221) ---- */
        for (uint32_t i = 0; i < n_uid; i++) if (!uid_file[uids[i]]) { uid_file[uids[i]] = 1; file_uid_count++; }
        for (int i = 0; i < 4; i++) f_lvl[i] |= lvl[i];
        for (int i = 0; i < 5; i++) f_svc[i] |= svc[i];
        for (int i = 0; i < 13; i++) { f_msg[i] |= wmsg[i]; f_tag[i] |= wtag[i]; }
        for (int i = 0; i < 8; i++) f_reg[i] |= reg[i];
        for (int i = 0; i < 3; i++) f_az[i] |= az[i];
        if ((bi + 1) % blocks_per_file == 0) {
            uint64_t fi = bi / blocks_per_file;
            uint64_t rows = (uint64_t)rows_per_block * blocks_per_file;
            uint64_t enums_tok = 0, enums_ft = 0;
            for (int i = 0; i < 4; i++) { enums_tok += f_lvl[i]; enums_ft += f_lvl[i]; }
            for (int i = 0; i < 5; i++) { enums_tok += f_svc[i]; enums_ft += f_svc[i]; }
            for (int i = 0; i < 13; i++) { enums_tok += (f_msg[i] | f_tag[i]); enums_ft += f_msg[i] + f_tag[i]; }
            for (int i = 0; i < 8; i++) { enums_tok += f_reg[i]; enums_ft += f_reg[i]; }
            for (int i = 0; i < 3; i++) { enums_tok += f_az[i]; enums_ft += f_az[i]; }
            c->file_counts[3 * fi + 0] = rows ? 9 : 0;
            c->file_counts[3 * fi + 1] = rows + file_uid_count + enums_tok;
            c->file_counts[3 * fi + 2] = rows + file_uid_count + enums_ft;
        }
    }
    c->group_begin[3 * n_blocks] = kb.nkeys;
    c->bytes = kb.bytes;
    c->key_off = kb.off;
    c->n_keys = kb.nkeys;
    free(uid_seen); free(uid_file); free(uids);
    return c;
}

/* accessors for ctypes */
const uint8_t *bgen_bytes(const bgen_corpus *c) { return c->bytes; }
const uint64_t *bgen_key_off(const bgen_corpus *c) { return c->key_off; }
uint64_t bgen_n_keys(const bgen_corpus *c) { return c->n_keys; }
const uint64_t *bgen_group_begin(const bgen_corpus *c) { return c->group_begin; }
const uint64_t *bgen_file_counts(const bgen_corpus *c) { return c->file_counts; }
uint64_t bgen_n_files(const bgen_corpus *c) { return c->n_files; }
