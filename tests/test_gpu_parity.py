"""-m gpu parity tests: the CUDA path (through the C ABI) against the CPU oracle,
bit for bit, on the same seeded inputs.  Integer / byte work: the bar is exact equality."""
from __future__ import annotations

import random

import numpy as np
import pytest

import bloomsearch_b200 as bs
from bloomsearch_b200 import _native as N
from oracle import bloomref as pyref
from oracle import cref
from oracle import murmur_canonical as canon
from synth.corpus import SynthCorpus
from tests.helpers import oracle_units, rand_keys, to_oracle_tuple

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------- K1 hash ---
def test_hash_keys_matches_oracle_all_lengths(ctx):
    rng = random.Random(1234)
    keys = [b""] + [bytes(rng.randrange(256) for _ in range(L)) for L in range(0, 100) for _ in range(8)]
    keys += [b"hello", b"hello, world", b"The quick brown fox jumps over the lazy dog.", b"service::auth"]
    got = ctx.hash_keys(keys)
    want = np.array([cref.base_hashes(k) for k in keys], dtype=np.uint64)
    assert np.array_equal(got, want)
    # public MurmurHash3_x64_128 vectors (h0,h1 are murmur(key, seed 0))
    assert tuple(int(x) for x in got[keys.index(b"hello")][:2]) == (0xCBD8A7B341BD9B02, 0x5B1E906A48AE1D19)


@pytest.mark.skipif(not canon.available(), reason="scikit-learn's copy of MurmurHash3.cpp is absent")
def test_hash_keys_match_canonical_murmurhash3(ctx):
    """THIRD-PARTY PIN: the device hash (csrc/bsg_device.cuh StreamHasher / base_hashes) against Austin Appleby's
    MurmurHash3.cpp compiled unmodified (oracle/_ref, oracle/murmur_canonical.py): h0,h1 = murmur(key),
    h2,h3 = murmur(key || 0x01) — what bloom/v3's sum256 documents — for every length 0..300, the virtual 0x01
    byte at tail offsets 0 / 8 / 15 after many blocks, and long keys."""
    rng = random.Random(77)
    lens = list(range(0, 301)) + [1000, 4096, 4097, 70001] + [16 * b + t for b in (7, 40, 200) for t in (0, 8, 15)]
    keys = [bytes(rng.randrange(256) for _ in range(L)) for L in lens for _ in range(2)] + [b"\x00" * 31, b"\xff" * 47, b"\x01"]
    got = ctx.hash_keys(keys)
    want = np.array([canon.base_hashes(k) for k in keys], dtype=np.uint64)
    assert np.array_equal(got, want)


def test_hash_keys_long_and_empty(ctx):
    rng = random.Random(5)
    keys = [bytes(rng.randrange(256) for _ in range(L)) for L in (0, 1, 15, 16, 17, 31, 32, 33, 255, 256, 4097, 70001)]
    got = ctx.hash_keys(keys)
    want = np.array([cref.base_hashes(k) for k in keys], dtype=np.uint64)
    assert np.array_equal(got, want)
    assert ctx.hash_keys([]).shape == (0, 4)


# ------------------------------------------------------------- K2/K3 build ---
def _build_case(ctx, groups, fpr, file_of=None, n_files=0, file_counts=None):
    """groups: list of key lists.  Returns (gpu_words, oracle_words)."""
    keys = [k for g in groups for k in g]
    blob, off = N.pack_keys(keys)
    group_begin = np.cumsum([0] + [len(g) for g in groups]).astype(np.uint64)
    desc, word_off = [], 0
    for g in groups:
        m, k = bs.estimate_parameters(max(len(g), 1), fpr)
        desc.append((m, k, word_off))
        word_off += (m + 63) // 64
    gf = np.arange(len(groups), dtype=np.uint32)
    gf2 = None
    if file_of is not None:
        fids = []
        for f in range(n_files):
            m, k = bs.estimate_parameters(max(file_counts[f], 1), fpr)
            fids.append(len(desc))
            desc.append((m, k, word_off))
            word_off += (m + 63) // 64
        gf2 = np.array([fids[f] if f >= 0 else N.NO_FILTER for f in file_of], dtype=np.uint32)
    d = np.array(desc, dtype=N.DESC_DTYPE)
    got = ctx.build(blob, off, group_begin, gf, gf2, d, word_off)
    want = cref.build_filters(blob, off, group_begin, gf, gf2, d, word_off)
    return got, want, d


def test_build_matches_oracle_small(ctx):
    rng = random.Random(7)
    groups = [rand_keys(rng, n, 1, 30) for n in (0, 1, 2, 9, 100, 101, 1000, 2030)]
    got, want, _ = _build_case(ctx, groups, 0.001)
    assert np.array_equal(got, want)


def test_build_sizing_matches_reference_pins(ctx):
    # file_format_test.go:28-94: counts {2,101,101} sized like NewWithEstimates(count, fpr)
    assert bs.estimate_parameters(2, 0.001) == cref.estimate_parameters(2, 0.001) == (29, 11)
    assert bs.estimate_parameters(101, 0.001) == (1453, 10)
    assert bs.estimate_parameters(2, 0.02) == (17, 6)  # lifecycle_durability_test.go:1195
    assert bs.estimate_parameters(100, 0.01) == (959, 7)  # bloom_tree_engine_test.go:366
    for n in (1, 3, 17, 1000, 10 ** 4, 10 ** 5, 10 ** 6, 12345678):
        for p in (0.5, 0.1, 0.01, 0.001, 1e-6, 1e-9):
            assert bs.estimate_parameters(n, p) == cref.estimate_parameters(n, p) == pyref.estimate_parameters(n, p)


def test_build_with_file_level_union(ctx):
    # flush.go:204-254: block filters + file filter sized for the union's distinct count
    c = SynthCorpus(42, 0, 20, 500, 10)
    groups, file_of = [], []
    for b in range(c.n_blocks):
        for kind in range(3):
            groups.append(c.group_keys(b, kind))
            file_of.append((b // c.blocks_per_file) * 3 + kind)
    counts = [int(c.file_counts[f // 3][f % 3]) for f in range(c.n_files * 3)]
    got, want, d = _build_case(ctx, groups, 0.001, file_of, c.n_files * 3, counts)
    assert np.array_equal(got, want)
    # and the file filter equals buildSizedBloomFilter(union) (duplicates are idempotent)
    n_block_filters = len(groups)
    for f in range(c.n_files * 3):
        union = set()
        for b in range((f // 3) * c.blocks_per_file, (f // 3 + 1) * c.blocks_per_file):
            union.update(c.group_keys(b, f % 3))
        assert len(union) == counts[f]
        ref = cref.Filter.build_sized(sorted(union), 0.001)
        m, k, wo = (int(x) for x in d[n_block_filters + f])
        assert (ref.m, ref.k) == (m, k)
        assert np.array_equal(got[wo:wo + ref.nwords], ref.words())


def test_build_large_filter_global_atomics_and_split_groups(ctx):
    # one filter of 300k keys: > shared-memory staging and > the 16k-key group split
    keys = [b"key-%d" % i for i in range(300_000)]
    got, want, d = _build_case(ctx, [keys], 0.001)
    assert int(d[0]["m"]) > 227 * 1024 * 8
    assert np.array_equal(got, want)


def test_build_shared_filter_across_groups(ctx):
    # shards of one entry set -> same filter: OR is idempotent and commutative
    rng = random.Random(3)
    keys = rand_keys(rng, 5000, 3, 20)
    blob, off = N.pack_keys(keys)
    m, k = bs.estimate_parameters(len(keys), 0.001)
    d = np.array([(m, k, 0)], dtype=N.DESC_DTYPE)
    gb = np.array([0, 1000, 1001, 3500, 5000], dtype=np.uint64)
    gf = np.zeros(4, dtype=np.uint32)
    got = ctx.build(blob, off, gb, gf, None, d, (m + 63) // 64)
    ref = cref.Filter.build_sized(keys, 0.001)
    assert np.array_equal(got, ref.words())


def test_build_edge_m_values(ctx):
    # Barrett reduction exactness on awkward moduli, incl. m == 1 and powers of two
    rng = random.Random(11)
    keys = rand_keys(rng, 64, 1, 12)
    blob, off = N.pack_keys(keys)
    ms = [1, 2, 3, 5, 63, 64, 65, 127, 128, 129, 4095, 4096, 4097, (1 << 20) - 1, 1 << 20, (1 << 20) + 1,
          (1 << 27) - 1, 1 << 27, (1 << 32) + 7]
    desc, wo = [], 0
    for i, m in enumerate(ms):
        desc.append((m, 1 + i % 13, wo))
        wo += (m + 63) // 64
    d = np.array(desc, dtype=N.DESC_DTYPE)
    n = len(ms)
    gb = (np.arange(0, n + 1) * 3).astype(np.uint64)  # 3 keys per filter, 57 <= 64 keys
    gf = np.arange(n, dtype=np.uint32)
    got = ctx.build(blob, off, gb, gf, None, d, wo)
    want = cref.build_filters(blob, off, gb, gf, None, d, wo)
    assert np.array_equal(got, want)


# --------------------------------------------------------------- K4 probe ---
def _probe_case(ctx, desc, words, keys, kinds, prog=None, paths=(N.PROBE_AUTO, N.PROBE_STAGED, N.PROBE_GATHER)):
    n_units = len(desc) // 3
    blob, off = N.pack_keys(keys)
    kinds = np.asarray(kinds, dtype=np.uint8)
    want_m = cref.probe_matrix(desc, words, n_units, blob, off, kinds)
    want_mask = cref.probe_mask(desc, words, n_units, blob, off, kinds, prog)
    corpus = bs.Corpus(ctx, desc, words)
    try:
        m, mask = corpus.probe(keys, kinds, prog)
        assert np.array_equal(m, want_m)
        assert np.array_equal(mask, want_mask)
        for path in paths:
            q = bs.Query(corpus, keys, kinds, prog)
            q.run(path)
            m2, mask2 = q.fetch()
            q.close()
            assert np.array_equal(m2, want_m), f"path {path}"
            assert np.array_equal(mask2, want_mask), f"path {path}"
    finally:
        corpus.close()
    return want_m, want_mask


def _mixed_keys(rng, unit_keys, n_present, n_absent):
    keys, kinds = [], []
    flat = [(kind, k) for kinds_ in unit_keys for kind, ks in enumerate(kinds_) for k in ks]
    for kind, k in rng.sample(flat, min(n_present, len(flat))):
        keys.append(k)
        kinds.append(kind)
    for i in range(n_absent):
        keys.append(b"absent%d" % i)
        kinds.append(i % 3)
    return keys, kinds


def test_probe_matrix_and_mask_small_corpus(ctx):
    rng = random.Random(21)
    unit_keys = [(rand_keys(rng, 9, 3, 12), rand_keys(rng, 200 + 13 * u, 1, 16), rand_keys(rng, 210 + 7 * u, 4, 30))
                 for u in range(37)]
    desc, words = oracle_units(unit_keys, 0.001)
    keys, kinds = _mixed_keys(rng, unit_keys, 60, 40)
    cq = bs.compile_bloom_query(bs.BloomQuery(bs.And(bs.Or(*[bs.Token(k) for k, kd in zip(keys, kinds) if kd == 1][:5]),
                                                     bs.Field(unit_keys[0][0][0]))))
    # use the raw key list with a hand-written program instead, exercising every op
    prog = np.array([(N.OP_LEAF, 0), (N.OP_LEAF, 1), (N.OP_OR, 2), (N.OP_LEAF, 2), (N.OP_TRUE, 0), (N.OP_AND, 3),
                     (N.OP_FALSE, 0), (N.OP_OR, 2)], dtype=N.OP_DTYPE)
    want_m, want_mask = _probe_case(ctx, desc, words, keys, kinds, prog)
    assert want_m.any() and not want_m.all()
    assert cq.prog is not None


@pytest.mark.parametrize("n_keys", [1, 2, 5, 8, 31, 32, 33, 64, 65, 511, 512, 513, 1000, 2049])
def test_probe_key_count_boundaries(ctx, n_keys):
    rng = random.Random(100 + n_keys)
    unit_keys = [(rand_keys(rng, 5, 3, 8), rand_keys(rng, 150, 1, 10), rand_keys(rng, 160, 4, 20)) for _ in range(19)]
    desc, words = oracle_units(unit_keys, 0.01)
    keys, kinds = _mixed_keys(rng, unit_keys, n_keys // 2, n_keys - n_keys // 2)
    _probe_case(ctx, desc, words, keys[:n_keys], kinds[:n_keys])


def test_probe_absent_filters_fail_open_and_empty_cases(ctx):
    rng = random.Random(9)
    unit_keys = [(rand_keys(rng, 4, 3, 8), rand_keys(rng, 50, 1, 10), rand_keys(rng, 50, 4, 20)) for _ in range(6)]
    absent = {(0, 0), (1, 1), (2, 2), (3, 0), (3, 1), (3, 2)}
    desc, words = oracle_units(unit_keys, 0.001, absent)
    keys, kinds = _mixed_keys(rng, unit_keys, 10, 11)
    want_m, _ = _probe_case(ctx, desc, words, keys, kinds)
    # unit 3 has no filters at all: every probe must say "maybe" (query_exec.go:137-151)
    assert bs.unpack_matrix(want_m, len(keys))[3].all()
    # no keys, no program: every unit survives (query_exec.go:81-83)
    corpus = bs.Corpus(ctx, desc, words)
    m, mask = corpus.probe([], [], None)
    assert bs.unpack_mask(mask, corpus.n_units).all()
    corpus.close()
    # empty corpus
    empty = bs.Corpus(ctx, np.zeros(0, N.DESC_DTYPE), np.zeros(0, np.uint64))
    m, mask = empty.probe([b"x"], [1], None)
    assert m.shape[0] == 0 and mask.shape[0] == 0
    empty.close()


def test_probe_large_filters_take_gather_path(ctx):
    # file-level sized filters (> shared memory): only the gather path can serve them
    big = [b"tok-%d" % i for i in range(400_000)]
    f_big = cref.Filter.build_sized(big, 0.001)
    small = [b"f%d" % i for i in range(9)]
    f_small = cref.Filter.build_sized(small, 0.001)
    desc = np.zeros(6, dtype=cref.DESC_DTYPE)
    wb, ws = f_big.words(), f_small.words()
    desc[0] = (f_small.m, f_small.k, 0)
    desc[1] = (f_big.m, f_big.k, len(ws))
    desc[3] = (f_small.m, f_small.k, 0)            # unit 1 shares the small field filter words
    desc[4] = (f_small.m, f_small.k, 0)
    words = np.concatenate([ws, wb])
    keys = [b"tok-5", b"tok-399999", b"nope", b"f3", b"f9"] + [b"absent%d" % i for i in range(70)]
    kinds = [1, 1, 1, 0, 0] + [1] * 70
    want_m, _ = _probe_case(ctx, desc, words, keys, kinds)
    bits = bs.unpack_matrix(want_m, len(keys))
    assert bits[0, 0] and bits[0, 1] and bits[0, 3] and not bits[0, 4]


def test_probe_big_endian_load_equals_native(ctx):
    rng = random.Random(77)
    unit_keys = [(rand_keys(rng, 5, 3, 8), rand_keys(rng, 300, 1, 10), rand_keys(rng, 300, 4, 20)) for _ in range(8)]
    desc, words = oracle_units(unit_keys, 0.001)
    keys, kinds = _mixed_keys(rng, unit_keys, 30, 30)
    blob, off = N.pack_keys(keys)
    want = cref.probe_matrix(desc, words, 8, blob, off, np.asarray(kinds, np.uint8))
    be = words.byteswap()  # what bitset.WriteTo puts on disk
    corpus = bs.Corpus(ctx, desc, be, big_endian=True)
    m, _ = corpus.probe(keys, kinds)
    corpus.close()
    assert np.array_equal(m, want)


def test_synth_corpus_config2_shape_parity(ctx):
    """BASELINE config 2 shape at a size the oracle finishes in seconds: GPU-built filters ==
    oracle-built filters, GPU candidate matrix == oracle matrix, staged == gather."""
    c = SynthCorpus(42, 0, 60, 1000, 10)
    counts = c.group_counts()
    desc = np.zeros(c.n_blocks * 3, dtype=N.DESC_DTYPE)
    wo = 0
    for g in range(c.n_blocks * 3):
        m, k = bs.estimate_parameters(max(int(counts.reshape(-1)[g]), 1), 0.001)
        desc[g] = (m, k, wo)
        wo += (m + 63) // 64
    gf = np.arange(c.n_blocks * 3, dtype=np.uint32)
    words = ctx.build(c.blob, c.key_off, c.group_begin, gf, None, desc, wo)
    want_words = cref.build_filters(c.blob, c.key_off, c.group_begin, gf, None, desc, wo, n_threads=4)
    assert np.array_equal(words, want_words)
    rng = random.Random(2)
    present = [c.key(i) for i in rng.sample(range(c.n_keys), 500)]
    kind_of = {}
    for b in range(c.n_blocks):
        for kind in range(3):
            for i in range(int(c.group_begin[3 * b + kind]), int(c.group_begin[3 * b + kind + 1])):
                kind_of.setdefault(i, kind)
    idx = rng.sample(range(c.n_keys), 500)
    keys = [c.key(i) for i in idx] + [b"absent%d" % i for i in range(500)]
    kinds = [kind_of[i] for i in idx] + [i % 3 for i in range(500)]
    _probe_case(ctx, desc, words, keys, kinds)
    assert present


# ------------------------------------------------- reference tests, mirrored ---
def test_evaluate_bloom_filters_reference_cases(ctx):
    """bloom_tree_engine_test.go:357-442 (TestEvaluateBloomFilters), same filters and queries,
    evaluated on the GPU and by the oracle's recursive evaluator."""
    def sized(entries):
        f = cref.Filter.with_estimates(100, 0.01)
        for e in entries:
            f.add(e)
        return f
    ff = sized([b"user.name", b"user.age"])
    tf = sized([b"alice", b"30"])
    ftf = sized([bs.make_field_token_key(b"user.name", b"alice"), bs.make_field_token_key(b"user.age", b"30")])
    assert (ff.m, ff.k) == (959, 7)
    unit = bs.BloomFilters(*(bs.BloomFilter(f.m, f.k, f.words()) for f in (ff, tf, ftf)))
    corpus = bs.Corpus.from_filters(ctx, [unit])
    cases = [
        ("nil query", None, True),
        ("field exists", bs.NewQuery().Field("user.name").Build(), True),
        ("field does not exist", bs.NewQuery().Field("nonexistent.field").Build(), False),
        ("token exists", bs.NewQuery().Token("alice").Build(), True),
        ("field-token exists", bs.NewQuery().FieldToken("user.name", "alice").Build(), True),
        ("OR one match", bs.NewQuery().Match(bs.Or(bs.Field("nonexistent.field"), bs.Field("user.name"))).Build(), True),
        ("AND one mismatch", bs.NewQuery().Match(bs.And(bs.Field("nonexistent.field"), bs.Field("user.name"))).Build(), False),
        ("OR field / field-token", bs.NewQuery().Match(bs.Or(bs.Field("nonexistent.field"),
                                                             bs.FieldToken("user.name", "alice"))).Build(), True),
    ]
    from tests.helpers import to_oracle_tuple
    for name, q, expected in cases:
        got = bool(corpus.evaluate_bloom_filters(q)[0])
        ref = cref.evaluate_bloom_filters(ff, tf, ftf, to_oracle_tuple(q.Expression) if q else None)
        assert got == ref == expected, name
    corpus.close()


def test_example_config1_roundtrip(ctx):
    """example_test.go:18-86 shape: 2 rows, 1 block; FieldToken("service","auth") keeps the block,
    a token that is not there prunes it."""
    es = bs.BloomEntrySets()
    rows = [{"id": 1, "service": "auth", "message": "login timeout for user"},
            {"id": 2, "service": "payment", "message": "charge succeeded"}]
    for row in rows:
        for path, value in row.items():
            p = path.encode()
            es.add_field(p)
            for tok in str(value).lower().split():
                es.add_token(tok.encode())
                es.add_field_token(p, tok.encode())
    filters = es.build_filters(ctx, 0.001)
    for got, entries in zip(filters.as_tuple(), (es.fields, es.tokens, es.fieldTokens)):
        ref = cref.Filter.build_sized(sorted(entries), 0.001)
        assert (got.m, got.k) == (ref.m, ref.k)
        assert np.array_equal(got.words, ref.words())
    corpus = bs.Corpus.from_filters(ctx, [filters])
    assert corpus.evaluate_bloom_filters(bs.NewQuery().FieldToken("service", "auth").Build())[0]
    assert not corpus.evaluate_bloom_filters(bs.NewQuery().FieldToken("service", "billing").Build())[0]
    assert not corpus.evaluate_bloom_filters(bs.NewQuery().Token("nonexistent-token").Build())[0]
    corpus.close()


# ------------------------------------------------ raw filter sections (A6/A7) ---
def test_load_sections_matches_oracle_decode(ctx):
    """bsg_corpus_load_sections == parseFilterSection per unit (file_format.go:392-448): framing,
    CRC32C, BE decode on the device; probe results equal the oracle's decode+probe."""
    rng = random.Random(31)
    unit_keys = [(rand_keys(rng, 3 + u % 7, 2, 9), rand_keys(rng, 100 + 37 * u, 1, 14), rand_keys(rng, 90 + 11 * u, 4, 25))
                 for u in range(23)]
    absent = {(1, 0), (2, 1), (3, 2), (4, 0), (4, 1), (4, 2)}
    desc, words = oracle_units(unit_keys, 0.001, absent)
    sec, sec_off = cref.encode_sections(desc, words, len(unit_keys))
    keys, kinds = _mixed_keys(rng, unit_keys, 40, 40)
    blob, off = N.pack_keys(keys)
    kinds = np.asarray(kinds, np.uint8)
    want, errs = cref.probe_sections_matrix(sec, sec_off, blob, off, kinds)
    assert errs == 0
    corpus, status = bs.Corpus.from_sections(ctx, sec, sec_off)
    assert corpus.n_bad == 0 and not status.any()
    for u in (0, 1, 4, 22):
        d = corpus.unit_desc(u)
        for k in range(3):
            assert (int(d[k]["m"]), int(d[k]["k"])) == (int(desc[u * 3 + k]["m"]), int(desc[u * 3 + k]["k"]) if desc[u * 3 + k]["m"] else 0)
    m, _ = corpus.probe(keys, kinds)
    corpus.close()
    assert np.array_equal(m, want)
    # unaligned section starts: shift everything by 3 bytes
    sec2 = np.concatenate([np.zeros(3, np.uint8), sec])
    off2 = sec_off.copy()
    off2[0] = 0
    sec2_off = np.concatenate([[np.uint64(3)], sec_off[1:] + np.uint64(3)]).astype(np.uint64)
    sec2_off = np.concatenate([[np.uint64(0)], sec2_off])  # unit 0 = the 3 junk bytes (too small -> error, kept)
    corpus, status = bs.Corpus.from_sections(ctx, sec2, sec2_off)
    assert status[0] == -1 and not status[1:].any() and corpus.n_bad == 1
    m2, _ = corpus.probe(keys, kinds)
    corpus.close()
    assert bs.unpack_matrix(m2, len(keys))[0].all()      # failed unit cannot be disqualified
    assert np.array_equal(m2[1:], want)


def test_load_sections_corruption_is_isolated_per_unit(ctx):
    # file_format_test.go:583-802 style corruption: byte flips, truncation, bad flags, trailing bytes
    rng = random.Random(32)
    unit_keys = [(rand_keys(rng, 4, 2, 9), rand_keys(rng, 60, 1, 14), rand_keys(rng, 60, 4, 25)) for _ in range(8)]
    desc, words = oracle_units(unit_keys, 0.001)
    secs = []
    for u in range(8):
        s1, _ = cref.encode_sections(desc[u * 3:u * 3 + 3], words, 1)
        secs.append(bytearray(s1.tobytes()))
    secs[1][len(secs[1]) // 2] ^= 0x10            # payload flip -> CRC mismatch (-2)
    secs[2][-1] ^= 0xFF                           # CRC field flip (-2)
    secs[3] = secs[3][:len(secs[3]) - 9]          # truncated (-2: CRC no longer matches)
    bad_flags = bytearray(secs[4]); bad_flags[0] |= 0x80
    import struct
    bad_flags[-4:] = struct.pack("<I", cref.crc32c(bytes(bad_flags[:-4])))
    secs[4] = bad_flags                           # valid CRC, unknown flag bit (-3)
    trailing = bytearray(secs[5][:-4]) + b"\x00\x01"
    trailing += struct.pack("<I", cref.crc32c(bytes(trailing)))
    secs[5] = trailing                            # valid CRC, trailing bytes (-7)
    blob_sec = np.frombuffer(b"".join(bytes(x) for x in secs), dtype=np.uint8)
    sec_off = np.cumsum([0] + [len(x) for x in secs]).astype(np.uint64)
    corpus, status = bs.Corpus.from_sections(ctx, blob_sec, sec_off)
    assert list(status) == [0, -2, -2, -2, -3, -7, 0, 0]
    assert corpus.n_bad == 5
    # the oracle (and the reference) reject the same sections
    for u in range(8):
        ok = True
        try:
            cref.section_parse(bytes(secs[u]))
        except ValueError:
            ok = False
        assert ok == (status[u] == 0)
    keys = [unit_keys[0][1][0], unit_keys[6][2][3], b"definitely-absent"]
    kinds = [1, 2, 1]
    m, _ = corpus.probe(keys, kinds)
    bits = bs.unpack_matrix(m, 3)
    corpus.close()
    assert bits[1:6].all()                        # failed units: fail open
    blob, off = N.pack_keys(keys)
    want = cref.probe_matrix(desc, words, 8, blob, off, np.asarray(kinds, np.uint8))
    wbits = bs.unpack_matrix(want, 3)
    for u in (0, 6, 7):
        assert np.array_equal(bits[u], wbits[u])
    # candidate mask: a block whose section failed to parse is an error, never a candidate
    # (query_exec.go:580-590 records the error and `continue`s) — with or without an expression
    corpus, status = bs.Corpus.from_sections(ctx, blob_sec, sec_off)
    q = bs.BloomQuery(bs.Or(bs.Token(keys[0]), bs.Token(unit_keys[3][1][5]), bs.Token(b"definitely-absent")))
    got = corpus.evaluate_bloom_filters(q)
    want_mask, errs = cref.probe_sections(blob_sec, sec_off, to_oracle_tuple(q.Expression))
    assert errs == 5 and np.array_equal(got, bs.unpack_mask(want_mask, 8)) and not got[1:6].any()
    got_all = corpus.evaluate_bloom_filters(None)
    assert list(got_all) == [True, False, False, False, False, False, True, True]
    corpus.close()
    # verify_crc=False accepts the payload-flipped section (like skipping the CRC would)
    corpus, status = bs.Corpus.from_sections(ctx, blob_sec, sec_off, verify_crc=False)
    assert status[1] == 0 and status[4] == -3
    corpus.close()


def test_bsg_probe_scratch_reuse_across_shapes(ctx):
    """bsg_probe keeps its device scratch between calls; pad words must never leak bits from a
    previous, differently shaped batch (1000 -> 970 -> 20 -> 1000 keys; with / without program)."""
    rng = random.Random(55)
    unit_keys = [(rand_keys(rng, 5, 3, 8), rand_keys(rng, 120, 1, 10), rand_keys(rng, 130, 4, 20)) for _ in range(41)]
    desc, words = oracle_units(unit_keys, 0.01)
    corpus = bs.Corpus(ctx, desc, words)
    all_keys, all_kinds = _mixed_keys(rng, unit_keys, 700, 400)
    for n in (1000, 970, 20, 1000, 33, 961):
        keys, kinds = all_keys[:n], np.asarray(all_kinds[:n], np.uint8)
        blob, off = N.pack_keys(keys)
        want = cref.probe_matrix(desc, words, len(unit_keys), blob, off, kinds)
        prog = np.array([(N.OP_LEAF, 0), (N.OP_LEAF, n - 1), (N.OP_OR, 2)], dtype=N.OP_DTYPE) if n % 2 else None
        want_mask = cref.probe_mask(desc, words, len(unit_keys), blob, off, kinds, prog)
        m, mask = corpus.probe(keys, kinds, prog)
        assert np.array_equal(m, want), n
        assert np.array_equal(mask, want_mask), n
    corpus.close()


def test_bsg_probe_matrix_only_rows_in_host_memory_odd_keys(ctx):
    """bsg_probe() without a mask: one upload, hashing fused into the staged kernel, matrix rows
    written by the kernel straight into pinned host memory.  Shapes change on the same scratch
    (pad words of a row must stay zero), keys include the empty key, keys with NUL bytes and a
    5 000-byte key; a corpus with one filter too large to stage falls back to the hash kernel."""
    rng = random.Random(77)
    odd = [b"", b"\x00", b"\x00\x00tail", bytes(rng.randrange(256) for _ in range(5000)), b"a" * 15, b"b" * 16, b"c" * 17]
    unit_keys = [(rand_keys(rng, 5, 3, 8), odd + rand_keys(rng, 120, 1, 10), rand_keys(rng, 130, 4, 20)) for _ in range(333)]
    desc, words = oracle_units(unit_keys, 0.001)
    corpus = bs.Corpus(ctx, desc, words)
    pool, pool_kinds = _mixed_keys(rng, unit_keys[:40], 900, 1300)
    pool = odd + pool
    pool_kinds = [1] * len(odd) + pool_kinds
    for n in (1000, 65, 1, 1025, 64, 2049, 33):
        keys, kinds = pool[:n], np.asarray(pool_kinds[:n], np.uint8)
        blob, off = N.pack_keys(keys)
        want = cref.probe_matrix(desc, words, len(unit_keys), blob, off, kinds)
        got = np.full((len(unit_keys), (n + 63) // 64), 0xDEADBEEFDEADBEEF, dtype=np.uint64)
        corpus.probe_packed(blob, off, kinds, None, got, None)
        assert np.array_equal(got, want), n
    bits = bs.unpack_matrix(want, 33)
    assert bits[:, :len(odd)].all()   # the odd keys are present in every unit
    corpus.close()
    # one unit whose token filter cannot be staged -> gather list is not empty -> separate hash kernel
    big = [b"tok-%d" % i for i in range(400_000)]
    f_big = cref.Filter.build_sized(big, 0.001)
    d2 = np.zeros(len(desc) + 3, dtype=cref.DESC_DTYPE)
    d2[:len(desc)] = desc
    d2[len(desc) + 1] = (f_big.m, f_big.k, len(words))
    w2 = np.concatenate([words, f_big.words()])
    c2 = bs.Corpus(ctx, d2, w2)
    keys = [b"tok-7", b"nope"] + pool[:100]
    kinds = np.asarray([1, 1] + pool_kinds[:100], np.uint8)
    blob, off = N.pack_keys(keys)
    want = cref.probe_matrix(d2, w2, len(d2) // 3, blob, off, kinds)
    got = np.zeros((len(d2) // 3, 2), dtype=np.uint64)
    c2.probe_packed(blob, off, kinds, None, got, None)
    assert np.array_equal(got, want)
    c2.close()


# ------------------------------------------------------ hierarchical probe ---
@pytest.mark.parametrize("n_extra_keys", [0, 60])
def test_hierarchical_probe_matches_two_stage_reference(ctx, n_extra_keys):
    """File stage (query_exec.go:399-406) then block stage only for blocks of surviving files
    (query_exec.go:572-615): block survives iff file filters pass AND its own filters pass."""
    c = SynthCorpus(11, 0, 48, 300, 6)   # 8 files x 6 blocks
    counts = c.group_counts().reshape(-1)
    bdesc = np.zeros(len(counts), dtype=N.DESC_DTYPE)
    wo = 0
    for g, n in enumerate(counts):
        m, k = bs.estimate_parameters(max(int(n), 1), 0.001)
        bdesc[g] = (m, k, wo)
        wo += (m + 63) // 64
    n_bw = wo
    fdesc = np.zeros(c.n_files * 3, dtype=N.DESC_DTYPE)
    for f in range(c.n_files):
        for kind in range(3):
            m, k = bs.estimate_parameters(max(int(c.file_counts[f][kind]), 1), 0.001)
            fdesc[f * 3 + kind] = (m, k, wo)
            wo += (m + 63) // 64
    gf = np.arange(len(bdesc), dtype=np.uint32)
    gf2 = np.array([len(bdesc) + (b // c.blocks_per_file) * 3 + kind for b in range(c.n_blocks) for kind in range(3)], np.uint32)
    words = ctx.build(c.blob, c.key_off, c.group_begin, gf, gf2, np.concatenate([bdesc, fdesc]), wo)
    want_words = cref.build_filters(c.blob, c.key_off, c.group_begin, gf, gf2, np.concatenate([bdesc, fdesc]), wo)
    assert np.array_equal(words, want_words)
    files = bs.Corpus(ctx, fdesc, words)      # descriptors index into the same words array
    blocks = bs.Corpus(ctx, bdesc, words)
    parent = (np.arange(c.n_blocks) // c.blocks_per_file).astype(np.uint32)
    blocks.set_parents(parent, c.n_files)
    ts_key = b"%d" % (1700000000 + 300 * 20 + 7)        # lives in block 20 (file 3) only
    uid = c.key(int(c.group_begin[3 * 40 + 1]) + 300 + 5)  # some user id token of block 40 (file 6)
    q = bs.BloomQuery(bs.Or(bs.Token(ts_key), bs.And(bs.Token(uid), bs.Field(b"nested.az")),
                            *[bs.Token(b"zz-absent-%d" % i) for i in range(n_extra_keys)]))
    got_f, got_b = bs.probe_hierarchical(files, blocks, q)
    cq = bs.compile_bloom_query(q)
    blob, off = N.pack_keys(cq.keys)
    wf = bs.unpack_mask(cref.probe_mask(fdesc, words, c.n_files, blob, off, cq.kinds, cq.prog), c.n_files)
    wb = bs.unpack_mask(cref.probe_mask(bdesc, words, c.n_blocks, blob, off, cq.kinds, cq.prog), c.n_blocks)
    assert np.array_equal(got_f, wf)
    assert np.array_equal(got_b, wb & wf[parent])
    assert got_b[20] and got_f[3] and 0 < got_f.sum() < c.n_files
    # no expression: everything survives both stages
    f_all, b_all = bs.probe_hierarchical(files, blocks, None)
    assert f_all.all() and b_all.all()
    # a corpus without parents is rejected
    with pytest.raises(bs.BloomGpuError):
        bs.probe_hierarchical(files, bs.Corpus(ctx, bdesc, words), q)
    files.close()
    blocks.close()
    assert n_bw > 0


def test_concurrent_callers_share_one_context(ctx):
    """The probe entry points are thread-safe and re-entrant (query_exec.go:303-357 runs up to
    MaxQueryConcurrency file workers): 8 host threads x 25 bsg_probe calls with different batches
    on one ctx / corpus, each checked against the oracle."""
    import threading
    rng = random.Random(77)
    unit_keys = [(rand_keys(rng, 6, 3, 8), rand_keys(rng, 180, 1, 10), rand_keys(rng, 170, 4, 20)) for _ in range(64)]
    desc, words = oracle_units(unit_keys, 0.001)
    corpus = bs.Corpus(ctx, desc, words)
    batches = []
    for t in range(8):
        keys, kinds = _mixed_keys(random.Random(1000 + t), unit_keys, 40 + 37 * t, 30 + 11 * t)
        kinds = np.asarray(kinds, np.uint8)
        blob, off = N.pack_keys(keys)
        prog = np.array([(N.OP_LEAF, 0), (N.OP_LEAF, 1), (N.OP_AND, 2), (N.OP_LEAF, len(keys) - 1), (N.OP_OR, 2)], dtype=N.OP_DTYPE)
        batches.append((keys, kinds, blob, off, prog,
                        cref.probe_matrix(desc, words, len(unit_keys), blob, off, kinds),
                        cref.probe_mask(desc, words, len(unit_keys), blob, off, kinds, prog)))
    errors = []

    def worker(t):
        keys, kinds, blob, off, prog, want_m, want_mask = batches[t]
        m = np.zeros_like(want_m)
        mask = np.zeros_like(want_mask)
        for _ in range(25):
            m[:] = 0
            mask[:] = 0
            corpus.probe_packed(blob, off, kinds, prog, m, mask)
            if not (np.array_equal(m, want_m) and np.array_equal(mask, want_mask)):
                errors.append(t)
                return

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    corpus.close()
    assert not errors, errors


# ------------------------------------------------------------ error behaviour ---
def test_invalid_arguments_are_rejected_with_codes(ctx):
    rng = random.Random(5)
    unit_keys = [(rand_keys(rng, 3, 3, 8), rand_keys(rng, 20, 1, 10), rand_keys(rng, 20, 4, 20)) for _ in range(3)]
    desc, words = oracle_units(unit_keys, 0.01)
    corpus = bs.Corpus(ctx, desc, words)
    keys, kinds = [b"a", b"b"], [1, 1]

    def bad(prog=None, kinds_=kinds):
        with pytest.raises(bs.BloomGpuError) as ei:
            corpus.probe(keys, kinds_, None if prog is None else np.array(prog, dtype=N.OP_DTYPE))
        assert ei.value.code == N.ERR_INVALID
        return ei.value.detail

    assert "leaf" in bad([(N.OP_LEAF, 2)])                       # leaf index out of range
    assert "pops" in bad([(N.OP_LEAF, 0), (N.OP_AND, 2)])         # stack underflow
    assert "stack" in bad([(N.OP_LEAF, 0), (N.OP_LEAF, 1)])       # leaves 2 values
    assert "unknown op" in bad([(9, 0)])
    assert "deeper" in bad([(N.OP_TRUE, 0)] * 65 + [(N.OP_AND, 65)])
    assert "kind" in bad(kinds_=[1, 3])
    corpus.close()
    # corpus descriptors
    for d, msg in (((1 << 63, 3, 0), "exceeds"), ((100, 0, 0), "k="), ((100, 3, 10**9), "outside")):
        dd = np.zeros(3, dtype=N.DESC_DTYPE)
        dd[1] = d
        with pytest.raises(bs.BloomGpuError) as ei:
            bs.Corpus(ctx, dd, np.zeros(4, np.uint64))
        assert ei.value.code == N.ERR_INVALID and msg in ei.value.detail
    # build: filter id out of range, non-monotone offsets
    blob, off = N.pack_keys([b"x", b"y"])
    dsc = np.array([(64, 3, 0)], dtype=N.DESC_DTYPE)
    with pytest.raises(bs.BloomGpuError):
        ctx.build(blob, off, np.array([0, 2], np.uint64), np.array([1], np.uint32), None, dsc, 1)
    badoff = off.copy()
    badoff[1] = 5
    with pytest.raises(bs.BloomGpuError):
        ctx.build(blob, badoff, np.array([0, 2], np.uint64), np.array([0], np.uint32), None, dsc, 1)
    with pytest.raises(bs.BloomGpuError):
        ctx.hash_keys_raw(blob, badoff) if hasattr(ctx, "hash_keys_raw") else (_ for _ in ()).throw(bs.BloomGpuError(-1, "n/a"))


# ------------------------------------------------- committed golden fixtures ---
def test_gpu_reproduces_oracle_pin_fixtures(ctx):
    """tests/golden/bloom_golden.json (oracle-generated, see make_golden.py) and the public
    murmur3 vectors, reproduced by the CUDA path alone: hashes, filter words, WriteTo bytes,
    the encoded section (host codec over GPU-built filters) and its device-side decode."""
    import json
    import os
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = json.load(open(os.path.join(gdir, "bloom_golden.json")))
    pub = json.load(open(os.path.join(gdir, "murmur3_x64_128.json")))["vectors"]
    keys = [e["key"].encode("utf-8") for e in g["base_hashes"]] + [v["data"].encode() for v in pub]
    h = ctx.hash_keys(keys)
    for i, e in enumerate(g["base_hashes"]):
        assert [int(x) for x in h[i]] == [int(x, 16) for x in e["h"]], e["key"]
    for j, v in enumerate(pub):
        row = h[len(g["base_hashes"]) + j]
        assert (int(row[0]), int(row[1])) == (int(v["h1"], 16), int(v["h2"], 16)), v["data"]
    for e in g["estimate_parameters"]:
        assert bs.estimate_parameters(e["n"], e["p"]) == (e["m"], e["k"])
    for e in g["filters"]:
        entries = [x.encode() for x in e["entries"]]
        blob, off = N.pack_keys(entries)
        d = np.array([(e["m"], e["k"], 0)], dtype=N.DESC_DTYPE)
        nw = (e["m"] + 63) // 64
        w = ctx.build(blob, off, np.array([0, len(entries)], np.uint64), np.zeros(1, np.uint32), None, d, nw)
        assert ["%016x" % int(x) for x in w] == e["words"], e["name"]
        assert bs.BloomFilter(e["m"], e["k"], w).write_to().hex() == e["write_to_hex"], e["name"]
    s = g["section"]
    es_f, es_t = bs.BloomEntrySets(), bs.BloomEntrySets()
    for x in s["field_entries"]:
        es_f.add_field(x.encode())
    for x in s["token_entries"]:
        es_t.add_token(x.encode())
    ff = es_f.build_filters(ctx, s["fpr"]).FieldBloomFilter
    tf = es_t.build_filters(ctx, s["fpr"]).TokenBloomFilter
    raw = bytes([3])
    for f in (ff, tf):
        body = f.write_to()
        raw += len(body).to_bytes(4, "little") + body
    raw += cref.crc32c(raw).to_bytes(4, "little")  # CRC by the checker; the framing is the product's
    assert raw.hex() == s["hex"]
    corpus, status = bs.Corpus.from_sections(ctx, np.frombuffer(bytes.fromhex(s["hex"]), dtype=np.uint8),
                                             np.array([0, len(s["hex"]) // 2], np.uint64))
    assert status[0] == 0
    m, _ = corpus.probe([b"service", b"nope", b"auth", b"zzz"], [0, 0, 1, 1])
    assert list(bs.unpack_matrix(m, 4)[0]) == [True, False, True, False]
    corpus.close()


# --------------------------------------------- fused field::token build (f.2) ---
def test_build_fieldtokens_equals_joined_keys(ctx):
    """bsg_build_fieldtokens(path, token pairs) == bsg_build / oracle on the materialised
    path + "::" + token keys (makeFieldTokenKey, tokenizer.go:508-511), all length / alignment mixes."""
    rng = random.Random(91)
    paths = [b"", b"a", b"level", b"nested.region", b"a.very.long.path.name.with.many.parts", b"x" * 15, b"y" * 16, b"z" * 17]
    tokens = [b"", b"1", b"info", b"region-7", b"0123456789abcde", b"0123456789abcdef", b"t" * 33] + \
             [bytes(rng.randrange(256) for _ in range(rng.randint(0, 40))) for _ in range(200)]
    strings = paths + tokens
    blob, off = N.pack_keys(strings)
    groups = []
    for g in range(5):
        n = [0, 1, 7, 300, 2000][g]
        groups.append([(rng.randrange(len(paths)), len(paths) + rng.randrange(len(tokens))) for _ in range(n)])
    pair_path = np.array([p for g in groups for p, _ in g], dtype=np.uint32)
    pair_token = np.array([t for g in groups for _, t in g], dtype=np.uint32)
    gb = np.cumsum([0] + [len(g) for g in groups]).astype(np.uint64)
    desc, wo = [], 0
    for g in groups:
        m, k = bs.estimate_parameters(max(len(set(g)), 1), 0.001)
        desc.append((m, k, wo))
        wo += (m + 63) // 64
    m, k = bs.estimate_parameters(sum(len(g) for g in groups), 0.001)   # a shared secondary (file-level) filter
    desc.append((m, k, wo))
    wo += (m + 63) // 64
    d = np.array(desc, dtype=N.DESC_DTYPE)
    gf = np.arange(len(groups), dtype=np.uint32)
    gf2 = np.full(len(groups), len(groups), dtype=np.uint32)
    got = ctx.build_fieldtokens(blob, off, pair_path, pair_token, gb, gf, gf2, d, wo)
    joined = [strings[p] + b"::" + strings[t] for g in groups for p, t in g]
    jb, jo = N.pack_keys(joined)
    want = cref.build_filters(jb, jo, gb, gf, gf2, d, wo)
    assert np.array_equal(got, want)
    assert np.array_equal(ctx.build(jb, jo, gb, gf, gf2, d, wo), want)
    with pytest.raises(bs.BloomGpuError):
        ctx.build_fieldtokens(blob, off, np.array([len(strings)], np.uint32), np.array([0], np.uint32),
                              np.array([0, 1], np.uint64), np.zeros(1, np.uint32), None, d[:1], wo)


# --------------------------------------------- exact distinct counts (f.3) ---
def _emissions(rng, n_groups, vocab, sizes):
    groups = [[vocab[rng.randrange(len(vocab))] for _ in range(sizes[g % len(sizes)])] for g in range(n_groups)]
    keys = [k for g in groups for k in g]
    gb = np.cumsum([0] + [len(g) for g in groups]).astype(np.uint64)
    return groups, keys, gb


def test_count_distinct_matches_host_sets(ctx):
    """bsg_count_distinct == len(set(...)) per group and per parent union — the counts the Go maps of
    bloomEntrySets hold (ingest.go:24-45 dedup, :105-123 unionInto/counts)."""
    rng = random.Random(1234)
    vocab = [b"", b"a", b"b", b"level", b"level\x00", b"service::api"] + \
            [bytes(rng.randrange(256) for _ in range(rng.randint(0, 48))) for _ in range(700)]
    groups, keys, gb = _emissions(rng, 23, vocab, [0, 1, 2, 31, 32, 33, 500, 4000])
    parent = np.array([g % 5 for g in range(len(groups))], dtype=np.uint32)
    blob, off = N.pack_keys(keys)
    gc, pc = ctx.count_distinct(blob, off, gb, parent, 6)       # parent 5 has no groups
    assert gc.tolist() == [len(set(g)) for g in groups]
    want_p = [len(set(k for g, p in zip(groups, parent) if p == q for k in g)) for q in range(6)]
    assert pc.tolist() == want_p and pc[5] == 0
    gc2, pc2 = ctx.count_distinct(blob, off, gb)
    assert gc2.tolist() == gc.tolist() and pc2 is None
    # no keys at all / only empty groups
    e_blob, e_off = N.pack_keys([])
    gc3, _ = ctx.count_distinct(e_blob, e_off, np.zeros(4, np.uint64))
    assert gc3.tolist() == [0, 0, 0]
    with pytest.raises(bs.BloomGpuError):
        ctx.count_distinct(blob, off, gb, np.full(len(groups), 9, np.uint32), 6)
    with pytest.raises(bs.BloomGpuError):
        ctx.count_distinct(blob, off, gb[:-1])                   # groups do not cover the keys


def test_counted_emissions_build_equals_deduped_sets(ctx):
    """Ingest without host maps: raw emissions (with repeats) -> device distinct counts -> sizes ->
    bsg_build over the raw emissions is bit-identical to buildFilters over the deduplicated sets
    (insertion is an idempotent OR; ingest.go:127-145), block filters and the file union filter."""
    rng = random.Random(77)
    vocab = [b"tok%d" % i for i in range(3000)]
    groups, keys, gb = _emissions(rng, 12, vocab, [50, 900, 2500, 1])
    blob, off = N.pack_keys(keys)
    parent = np.zeros(len(groups), dtype=np.uint32)
    gc, pc = ctx.count_distinct(blob, off, gb, parent, 1)
    desc, wo = [], 0
    for n in list(gc) + [pc[0]]:
        m, k = bs.estimate_parameters(max(int(n), 1), 0.001)
        desc.append((m, k, wo))
        wo += (m + 63) // 64
    d = np.array(desc, dtype=N.DESC_DTYPE)
    gf = np.arange(len(groups), dtype=np.uint32)
    gf2 = np.full(len(groups), len(groups), dtype=np.uint32)
    got = ctx.build(blob, off, gb, gf, gf2, d, wo)
    dedup = [sorted(set(g)) for g in groups]
    dk = [k for g in dedup for k in g]
    db, do = N.pack_keys(dk)
    dgb = np.cumsum([0] + [len(g) for g in dedup]).astype(np.uint64)
    # the oracle sizes from its own set sizes: same descriptors must come out
    d_want = []
    wo2 = 0
    for n in [len(g) for g in dedup] + [len(set(keys))]:
        m, k = pyref.estimate_parameters(max(n, 1), 0.001)
        d_want.append((m, k, wo2))
        wo2 += (m + 63) // 64
    assert d_want == desc
    want = cref.build_filters(db, do, dgb, gf, gf2, d, wo)
    assert np.array_equal(got, want)


# ---------------------------------------- staged kernels: every variant, forced refills ---
def _units_with_fprs(unit_keys, fprs, absent=()):
    """Like helpers.oracle_units but with one fpr per unit (k from 1 to 30)."""
    desc = np.zeros(len(unit_keys) * 3, dtype=cref.DESC_DTYPE)
    chunks, off = [], 0
    for u, kinds_ in enumerate(unit_keys):
        for kind, ks in enumerate(kinds_):
            if (u, kind) in absent:
                continue
            f = cref.Filter.build_sized(ks, fprs[u % len(fprs)])
            w = f.words()
            desc[u * 3 + kind] = (f.m, f.k, off)
            chunks.append(w)
            off += len(w)
    return desc, np.concatenate(chunks)


@pytest.mark.parametrize("variant,stages,n_units", [
    (0, 0, 700), (0, 2, 700), (1, 0, 700), (1, 2, 700), (1, 1, 700), (2, 0, 700), (2, 3, 700),
    # shapes whose B warps work in teams: rings that are a multiple of the team count (teams on, with
    # refills), rings that are not (rounded down / one-team fallback)
    (3, 0, 700), (3, 4, 700), (3, 1, 700), (3, 3, 700), (3, 7, 1300), (4, 0, 700), (4, 8, 2600), (4, 2, 700),
    (5, 0, 700), (5, 4, 700), (5, 3, 700)])
def test_probe_staged_variants_low_k_full_queue_and_refills(variant, stages, n_units, monkeypatch):
    """probe_staged (one phase) and every shape of probe_staged2 (phase A: locations 0..NT-1 of every
    key, phase B: locations NT..k-1 of the compacted survivors) must give the oracle's matrix for
    k = 1..5 (k <= NT: no phase-B work), k = 30, absent filters, keys present in every unit (survivor queue full),
    1024-key passes + a ragged second pass, and rings of 1-3 stages so every CTA refills."""
    from tests.conftest import _has_gpu
    if not _has_gpu():
        pytest.skip("no CUDA device in this process")
    monkeypatch.setenv("BSG_PROBE_VARIANT", str(variant))
    if stages:
        monkeypatch.setenv("BSG_PROBE_STAGES", str(stages))
    rng = random.Random(4242 + variant)
    shared_tok = rand_keys(rng, 700, 2, 14)       # present in every unit
    shared_ft = rand_keys(rng, 500, 6, 24)
    unit_keys = []
    for u in range(n_units):
        unit_keys.append((rand_keys(rng, 6, 3, 9), shared_tok + rand_keys(rng, 40 + u % 50, 15, 20),
                          shared_ft + rand_keys(rng, 30 + u % 40, 25, 32)))
    fprs = [0.6, 0.3, 0.2, 0.1, 0.05, 0.001, 1e-9]
    absent = {(5, 1), (6, 2), (7, 0), (8, 0), (8, 1), (8, 2), (n_units - 1, 1)}
    desc, words = _units_with_fprs(unit_keys, fprs, absent)
    assert sorted({int(k) for k in desc["k"] if k})[:5] == [1, 2, 3, 4, 5] and int(desc["k"].max()) == 30
    keys = shared_tok + shared_ft[:324]            # 1024 keys that pass everywhere: queue = 1024 survivors
    kinds = [1] * len(shared_tok) + [2] * 324
    extra, extra_kinds = _mixed_keys(rng, unit_keys[:50], 300, 401)  # present in few units + absent
    keys, kinds = keys + extra, kinds + extra_kinds
    assert len(keys) == 1725
    blob, off = N.pack_keys(keys)
    kinds = np.asarray(kinds, dtype=np.uint8)
    want = cref.probe_matrix(desc, words, n_units, blob, off, kinds)
    c2 = bs.Context(0)
    try:
        corpus = bs.Corpus(c2, desc, words)
        for _ in range(2):                         # twice: stage state must be clean at kernel exit
            q = bs.Query(corpus, keys, kinds, None)
            q.run(N.PROBE_STAGED)
            got, _ = q.fetch()
            q.close()
            assert np.array_equal(got, want), f"variant {variant} stages {stages} units {n_units}"
        # first 1024 keys pass in every unit that has the filter (no false negatives)
        bits = bs.unpack_matrix(want, len(keys))
        assert bits[:, :1024].all()
        corpus.close()
    finally:
        c2.close()


# ------------------------------- tile ring (probe_tiles_kernel): shapes, modes, masked fills ---
def _tiles_case(monkeypatch, env, n_units, key_kinds=(0, 1, 2), big=False, n_keys_cap=None):
    """probe_tiles_kernel against the oracle for one (shape, mode, grouping, ring length) setting:
    k = 1..5 and 30, absent filters, keys present in every unit (every key survives phase A), a ragged
    second pass, and — key_kinds — batches that touch only some filter kinds (masked stage fills)."""
    from tests.conftest import _has_gpu
    if not _has_gpu():
        pytest.skip("no CUDA device in this process")
    monkeypatch.setenv("BSG_PROBE_VARIANT", "6")
    for k, v in env.items():
        monkeypatch.setenv(k, str(v))
    rng = random.Random(777 + n_units + len(env))
    shared_tok = rand_keys(rng, 700, 2, 14)       # present in every unit
    shared_ft = rand_keys(rng, 500, 6, 24)
    unit_keys = []
    for u in range(n_units):
        extra_t = 40 + u % 50 + (9000 if big and u % 3 == 0 else 0)
        extra_f = 30 + u % 40 + (14000 if big and u % 5 == 0 else 0)
        unit_keys.append((rand_keys(rng, 6, 3, 9), shared_tok + rand_keys(rng, extra_t, 15, 20),
                          shared_ft + rand_keys(rng, extra_f, 25, 32)))
    fprs = [0.6, 0.3, 0.2, 0.1, 0.05, 0.001, 1e-9]
    absent = {(5, 1), (6, 2), (7, 0), (8, 0), (8, 1), (8, 2), (n_units - 1, 1)}
    desc, words = _units_with_fprs(unit_keys, fprs, absent)
    keys = shared_tok + shared_ft[:324]
    kinds = [1] * len(shared_tok) + [2] * 324
    extra, extra_kinds = _mixed_keys(rng, unit_keys[:50], 300, 401)
    keys, kinds = keys + extra, kinds + extra_kinds
    sel = [i for i, kd in enumerate(kinds) if kd in key_kinds]
    if n_keys_cap:
        sel = sel[:n_keys_cap]
    keys, kinds = [keys[i] for i in sel], [kinds[i] for i in sel]
    blob, off = N.pack_keys(keys)
    kinds = np.asarray(kinds, dtype=np.uint8)
    want = cref.probe_matrix(desc, words, n_units, blob, off, kinds)
    c2 = bs.Context(0)
    try:
        corpus = bs.Corpus(c2, desc, words)
        for _ in range(2):                         # twice: no state may leak between launches
            q = bs.Query(corpus, keys, kinds, None)
            q.run(N.PROBE_STAGED)
            got, _ = q.fetch()
            q.close()
            assert np.array_equal(got, want), f"env {env} units {n_units} kinds {key_kinds}"
        m, _ = corpus.probe(keys, kinds, None, want_mask=False)   # bsg_probe: hashing fused into the kernel
        assert np.array_equal(m, want), f"bsg_probe env {env}"
        corpus.close()
    finally:
        c2.close()


@pytest.mark.parametrize("shape", range(6))
def test_probe_tiles_every_shape(shape, monkeypatch):
    _tiles_case(monkeypatch, {"BSG_TILES_SHAPE": shape}, 700)


@pytest.mark.parametrize("env,n_units", [
    ({"BSG_TILE_MODE": 1, "BSG_TILE_UNITS": 1}, 700),                          # UNIT mode, one unit per tile
    ({"BSG_TILE_MODE": 1, "BSG_TILE_UNITS": 3, "BSG_PROBE_STAGES": 2}, 700),   # ragged groups, 2-stage ring
    ({"BSG_TILE_MODE": 1, "BSG_TILE_UNITS": 8, "BSG_TILE_BYTES": 200000}, 1300),
    ({"BSG_TILE_MODE": 1, "BSG_PROBE_STAGES": 1}, 700),                        # ring of one stage
    ({"BSG_TILE_MODE": 2}, 700),                                               # KIND mode forced on small units
    ({"BSG_TILE_MODE": 2, "BSG_PROBE_STAGES": 3, "BSG_TILES_SHAPE": 1}, 700),
    ({"BSG_TILE_MODE": 1, "BSG_TILE_UNITS": 1, "BSG_TILE_MIN_STAGES": 2, "BSG_TILES_SHAPE": 2}, 300),
    ({"BSG_TILE_MODE": 2, "BSG_PROBE_STAGES": 1}, 450),
    ({"BSG_PROBE_PDL": 0}, 300),
    ({"BSG_TILES_SHAPE": 3, "BSG_PROBE_STAGES": 4}, 900),
    ({"BSG_TILES_SHAPE": 0, "BSG_TILE_MODE": 2, "BSG_PROBE_STAGES": 5}, 450),
    ({"BSG_TILES_SHAPE": 4, "BSG_TILE_MODE": 1, "BSG_TILE_UNITS": 3, "BSG_TILE_BYTES": 9000}, 2000),
])
def test_probe_tiles_modes_groupings_rings(env, n_units, monkeypatch):
    _tiles_case(monkeypatch, env, n_units)


@pytest.mark.parametrize("env,cap", [
    ({"BSG_TILE_MODE": 1, "BSG_TILE_UNITS": 2}, 600),    # 600 keys present in both units of a tile: |L1| = 1200
    ({"BSG_TILE_MODE": 1, "BSG_TILE_UNITS": 3}, 400),    # 3 x 400 = 1200
    ({"BSG_TILE_MODE": 1, "BSG_TILE_UNITS": 5, "BSG_TILES_SHAPE": 0}, 120),   # 512-thread shape: 600 = 512 + 88
])
def test_probe_tiles_survivor_list_just_above_a_full_pass(env, cap, monkeypatch):
    """|L1| just above a multiple of the CTA size (a short, ragged last pass of round B1), with keys that pass
    every location, absent filters and k = 1..30."""
    _tiles_case(monkeypatch, env, 300, key_kinds=(1,), n_keys_cap=cap)


@pytest.mark.parametrize("env,key_kinds", [
    ({"BSG_TILE_MODE": 1}, (1,)), ({"BSG_TILE_MODE": 1}, (0, 2)), ({"BSG_TILE_MODE": 1, "BSG_TILE_UNITS": 5}, (2,)),
    ({"BSG_TILE_MODE": 2}, (1,)), ({"BSG_TILE_MODE": 2}, (2,)), ({"BSG_TILE_MODE": 2}, (0,)), ({"BSG_TILE_MODE": 2}, (0, 2)),
    ({"BSG_TILE_MODE": 1, "BSG_TILES_SHAPE": 1}, (1,)), ({"BSG_TILE_MODE": 2, "BSG_TILES_SHAPE": 1}, (0, 2)),
])
def test_probe_tiles_masked_fills(env, key_kinds, monkeypatch):
    """A batch that touches only some kinds copies only those filters into the stages."""
    _tiles_case(monkeypatch, env, 500, key_kinds=key_kinds)


def test_probe_tiles_large_units_kind_mode_and_gather_mix(monkeypatch):
    """Units of ~20-45 KB mixed with small ones: the corpus picks its mode by itself; small batches
    still work (most A warps hold no key)."""
    _tiles_case(monkeypatch, {}, 160, big=True)
    _tiles_case(monkeypatch, {}, 90, big=True, n_keys_cap=7)


@pytest.mark.parametrize("unit_kb", [100, 150, 400])
def test_probe_units_beyond_the_old_75kb_limit(ctx, unit_kb):
    """Units of 100 KB and 150 KB are staged per kind (probe_tiles KIND mode); 400 KB units exceed
    shared memory and take the gather path.  All must equal the oracle."""
    rng = random.Random(unit_kb)
    n_tok = unit_kb * 1024 * 8 // 2 // 15          # ~14.4 bits per key at fpr 0.001, two big filters per unit
    unit_keys = []
    for u in range(5):
        toks = [b"t%d-%d" % (u, i) for i in range(n_tok)]
        fts = [b"f::t%d-%d" % (u, i) for i in range(n_tok)]
        unit_keys.append(([b"f%d" % i for i in range(9)], toks, fts))
    desc, words = oracle_units(unit_keys, 0.001)
    keys, kinds = [], []
    for u in range(5):
        for i in rng.sample(range(n_tok), 40):
            keys += [b"t%d-%d" % (u, i), b"f::t%d-%d" % (u, i)]
            kinds += [1, 2]
    for i in range(300):
        keys.append(b"absent%d" % i)
        kinds.append(i % 3)
    _probe_case(ctx, desc, words, keys, kinds)


# ------------------------------------------------ resident filter cache (§8 f.4) ---
def test_filter_cache_load_query_invalidate_evict(ctx):
    """load -> query -> merge-invalidate -> query, LRU eviction under a byte budget, pins that outlive an
    invalidation; every answer equals the oracle's for the file's CURRENT filters."""
    rng = random.Random(88)

    def make_file(seed, n_units):
        r = random.Random(seed)
        uk = [(rand_keys(r, 4, 2, 9), rand_keys(r, 80, 1, 14), rand_keys(r, 80, 4, 25)) for _ in range(n_units)]
        d, w = oracle_units(uk, 0.001)
        sec, so = cref.encode_sections(d, w, n_units)
        return uk, d, w, sec, so
    files = {fid: make_file(1000 + fid, 6) for fid in range(5)}
    one = bs.Corpus.from_sections(ctx, files[0][3], files[0][4])[0]
    per_file = one.device_bytes()
    one.close()
    cache = bs.FilterCache(ctx, int(per_file * 3.5))     # room for three files

    def query_file(fid, version=None):
        uk, d, w, sec, so = version or files[fid]
        cp = cache.acquire(fid)
        hit = cp is not None
        if cp is None:                                    # miss: the host reads the sections and inserts them
            cp, status = cache.insert_sections(fid, sec, so)
            assert not status.any()
        key = uk[2][1][3]
        q = bs.BloomQuery(bs.Or(bs.Token(key), bs.Token(b"definitely-absent")))
        got = cp.evaluate_bloom_filters(q)
        cache.release(cp)
        want, errs = cref.probe_sections(sec, so, to_oracle_tuple(q.Expression))
        assert errs == 0 and np.array_equal(got, bs.unpack_mask(want, len(so) - 1)) and got[2]
        return hit
    assert [query_file(f) for f in (0, 1, 2)] == [False, False, False]
    assert [query_file(f) for f in (0, 1, 2)] == [True, True, True]
    st = cache.stats()
    assert st["entries"] == 3 and st["hits"] == 3 and st["misses"] == 3 and st["evictions"] == 0
    assert query_file(3) is False                         # over budget: the least recently used file (0) goes
    st = cache.stats()
    assert st["entries"] == 3 and st["evictions"] == 1 and st["used_bytes"] <= per_file * 3.5
    assert query_file(1) is True and query_file(0) is False
    # a merge rewrites file 1 (new filters under the same id): the old corpus must never be served again, but a
    # query that still holds it may finish
    held = cache.acquire(1)
    cache.invalidate(1)
    assert cache.acquire(1) is None
    newer = make_file(2001, 6)
    assert query_file(1, newer) is False                  # reload: answers come from the NEW filters
    old_key = files[1][0][2][1][3]
    got_old = held.evaluate_bloom_filters(bs.BloomQuery(bs.Token(old_key)))
    assert got_old[2]                                     # the pinned old corpus is still intact
    cache.release(held)
    files[1] = newer
    assert query_file(1) is True
    assert cache.stats()["invalidations"] == 1
    cache.close()


def test_bsg_probe_into_pinned_caller_buffer(ctx):
    """out_matrix in a bsg_host_alloc buffer: the kernels write the rows (pad words included) straight into it;
    reusing the buffer with another key count must leave no stale bits."""
    rng = random.Random(61)
    unit_keys = [(rand_keys(rng, 5, 3, 8), rand_keys(rng, 900, 1, 10), rand_keys(rng, 900, 4, 20)) for _ in range(300)]
    desc, words = oracle_units(unit_keys, 0.001)
    corpus = bs.Corpus(ctx, desc, words)
    buf = ctx.host_alloc((300, 16), np.uint64)
    for n_keys in (1000, 930, 97, 1000):
        keys, kinds = _mixed_keys(rng, unit_keys, n_keys // 2, n_keys - n_keys // 2)
        blob, off = N.pack_keys(keys)
        kinds = np.asarray(kinds, np.uint8)
        want = cref.probe_matrix(desc, words, 300, blob, off, kinds)
        out = buf[:, :(n_keys + 63) // 64] if (n_keys + 63) // 64 == 16 else np.zeros((300, (n_keys + 63) // 64), np.uint64)
        corpus.probe_packed(blob, off, kinds, None, out, None)
        assert np.array_equal(out, want), f"n_keys {n_keys}"
    corpus.close()
    ctx.host_free(buf)


# ------------------------------------------------- query batching (§8 f.4) ---
def _random_queries(rng, unit_keys, n):
    """Small BloomQueries in the shapes the reference's tests use (bloom_tree_engine_test.go:357-442): single
    conditions, AND / OR of a few conditions with present and absent keys, nested, nil."""
    def cond():
        u = rng.randrange(len(unit_keys))
        kind = rng.randrange(3)
        pool = unit_keys[u][kind]
        key = rng.choice(pool) if (pool and rng.random() < 0.6) else b"nope%d" % rng.randrange(1000)
        if kind == 0:
            return bs.Field(key)
        if kind == 1:
            return bs.Token(key)
        f, _, t = key.partition(b"::")     # the fixture's fieldtoken keys are field + "::" + token (tokenizer.go:508-511)
        return bs.FieldToken(f, t)
    out = []
    for i in range(n):
        r = rng.random()
        if r < 0.05:
            out.append(None)
        elif r < 0.3:
            out.append(bs.BloomQuery(cond()))
        elif r < 0.6:
            out.append(bs.BloomQuery(bs.And(*[cond() for _ in range(rng.randint(2, 4))])))
        elif r < 0.85:
            out.append(bs.BloomQuery(bs.Or(*[cond() for _ in range(rng.randint(2, 5))])))
        else:
            out.append(bs.BloomQuery(bs.And(bs.Or(cond(), cond()), bs.Or(cond(), cond(), cond()), cond())))
    return out


def _multi_fixture(ctx, n_units=96, seed=31, corrupt=False):
    rng = random.Random(seed)
    unit_keys = [(rand_keys(rng, 5, 3, 8), rand_keys(rng, 120, 1, 10),
                  sorted({a + b"::" + b for a, b in zip(rand_keys(rng, 110, 2, 8), rand_keys(rng, 110, 1, 9))}))
                 for _ in range(n_units)]
    desc, words = oracle_units(unit_keys, 0.01, absent={(3, 1), (7, 0), (7, 1), (7, 2)})
    return rng, unit_keys, bs.Corpus(ctx, desc, words)


def test_probe_multi_equals_separate_probes(ctx, monkeypatch):
    """bsg_probe_multi: one pass for the union of the keys, one mask per query == n separate bsg_probe calls
    (each of which is oracle-checked elsewhere), on both staged kernels and the gather path."""
    for env in ({}, {"BSG_PROBE_VARIANT": "3"}, {"BSG_PROBE_VARIANT": "6"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        c2 = bs.Context(0)
        rng, unit_keys, corpus = _multi_fixture(c2)
        for n in (1, 7, 130):   # 130 queries: > 1 024 keys in total -> two key passes
            queries = _random_queries(rng, unit_keys, n)
            got = corpus.probe_multi(queries)
            want = np.stack([corpus.evaluate_bloom_filters(q) for q in queries])
            assert got.shape == want.shape and np.array_equal(got, want), (env, n)
        # only nil queries: no keys at all, every unit survives
        assert corpus.probe_multi([None, None]).all()
        corpus.close()
        c2.close()
        for k in env:
            monkeypatch.delenv(k)


def test_probe_multi_rejects_bad_programs(ctx):
    rng, unit_keys, corpus = _multi_fixture(ctx, n_units=8)
    blob, off = N.pack_keys([b"a", b"b"])
    kinds = np.array([1, 1], np.uint8)
    qbegin = np.array([0, 1, 2], np.uint32)
    out = np.zeros((2, 1), np.uint64)
    bad = np.array([(N.OP_LEAF, 0), (N.OP_LEAF, 1), (N.OP_AND, 2)], dtype=N.OP_DTYPE)   # query 0 has ONE key: leaf 1 is out of range
    pbegin = np.array([0, 3, 3], np.uint32)
    rc = N.lib().bsg_probe_multi(ctx.handle, corpus.handle, N.ptr(blob), N.ptr(off), 2, N.ptr(kinds), 2, N.ptr(qbegin),
                                 N.ptr(bad), N.ptr(pbegin), N.ptr(out))
    assert rc == N.ERR_INVALID
    qbad = np.array([0, 2, 1], np.uint32)
    rc = N.lib().bsg_probe_multi(ctx.handle, corpus.handle, N.ptr(blob), N.ptr(off), 2, N.ptr(kinds), 2, N.ptr(qbad),
                                 N.ptr(bad), N.ptr(pbegin), N.ptr(out))
    assert rc == N.ERR_INVALID
    corpus.close()


@pytest.mark.parametrize("window_us,max_keys", [(0, 0), (200, 0), (0, 16)])
def test_batcher_concurrent_queries_share_launches(ctx, window_us, max_keys):
    """bsg_batcher: 16 threads x 30 queries each; every caller gets exactly its own bsg_probe mask, and the
    batcher needed fewer launches than calls (concurrent queries were merged)."""
    import threading
    rng, unit_keys, corpus = _multi_fixture(ctx, n_units=200, seed=5)
    n_threads, per_thread = 16, 30
    queries = [_random_queries(random.Random(100 + t), unit_keys, per_thread) for t in range(n_threads)]
    want = [[corpus.evaluate_bloom_filters(q) for q in qs] for qs in queries]
    batcher = bs.Batcher(corpus, max_keys=max_keys, window_us=window_us)
    errors = []
    barrier = threading.Barrier(n_threads)

    def worker(t):
        barrier.wait()
        for i, q in enumerate(queries[t]):
            got = batcher.evaluate(q)
            if not np.array_equal(got, want[t][i]):
                errors.append((t, i))
                return

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    st = batcher.stats()
    batcher.close()
    corpus.close()
    assert not errors, errors[:5]
    assert st["calls"] == n_threads * per_thread
    assert st["launches"] + st["bypassed"] <= st["calls"]
    if window_us:   # with a window the callers certainly met (without one, short Python-driven calls may never overlap)
        assert st["launches"] < st["calls"], st
        assert st["largest_batch"] >= 2, st


def test_batcher_bad_member_fails_alone(ctx):
    import threading
    rng, unit_keys, corpus = _multi_fixture(ctx, n_units=40, seed=9)
    batcher = bs.Batcher(corpus, window_us=20000)     # a long window so the two callers share a batch
    good = bs.BloomQuery(bs.Token(unit_keys[0][1][0]))
    want = corpus.evaluate_bloom_filters(good)
    res = {}

    def ok():
        res["good"] = batcher.evaluate(good)

    def bad():
        blob, off = N.pack_keys([b"x"])
        prog = np.array([(N.OP_LEAF, 5)], dtype=N.OP_DTYPE)   # leaf out of range
        try:
            batcher.evaluate_packed(blob, off, np.array([1], np.uint8), prog, np.zeros(1, np.uint64))
            res["bad"] = "no error"
        except bs.BloomGpuError as e:
            res["bad"] = e.code
    ths = [threading.Thread(target=ok), threading.Thread(target=bad)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    batcher.close()
    corpus.close()
    assert res["bad"] == N.ERR_INVALID
    assert np.array_equal(res["good"], want)


def test_gather_short_circuit_mask_equals_oracle(ctx):
    """Mask-only small queries on the gather path stop testing a unit's keys once its expression is decided (the
    batched form of evaluateBloomExpression's short circuit, query_exec.go:105-119): the mask must still be exactly
    the oracle's, for nested AND / OR trees, TRUE / FALSE ops, absent filters, 1..32 keys, with and without the
    short circuit, and in the hierarchical call (pruned parents)."""
    rng, unit_keys, corpus = _multi_fixture(ctx, n_units=300, seed=77)
    # the oracle needs the fixture's descriptors and words: rebuild them the way the fixture did
    desc2, words = oracle_units(unit_keys, 0.01, absent={(3, 1), (7, 0), (7, 1), (7, 2)})
    queries = _random_queries(rng, unit_keys, 60)
    # a wide one: 32 leaves under nested ORs inside an AND, plus constants
    wide = bs.BloomQuery(bs.And(bs.Or(*[bs.Token(k) for k in unit_keys[0][1][:15]]),
                                bs.Or(*[bs.Token(b"zz%d" % i) for i in range(10)], bs.Token(unit_keys[5][1][0])),
                                bs.Or(*[bs.FieldToken(*k.split(b"::", 1)) for k in unit_keys[9][2][:6]])))
    queries.append(wide)
    n_sc = 0
    for qy in queries:
        cq = bs.compile_bloom_query(qy)
        if cq.prog is None or len(cq.keys) == 0:
            continue
        blob, off = N.pack_keys(cq.keys)
        want = cref.probe_mask(desc2, words, corpus.n_units, blob, off, np.asarray(cq.kinds, np.uint8), cq.prog)
        q = bs.Query(corpus, cq.keys, cq.kinds, cq.prog)
        for path in (N.PROBE_GATHER, N.PROBE_AUTO):
            q.run(path, want_matrix=False)
            _, mask = q.fetch(want_matrix=False)
            assert np.array_equal(mask, want), (path, cq.keys[:3])
        q.run(N.PROBE_GATHER, want_matrix=True)
        m_exact, mask2 = q.fetch()
        assert np.array_equal(mask2, want)
        assert np.array_equal(m_exact, cref.probe_matrix(desc2, words, corpus.n_units, blob, off, np.asarray(cq.kinds, np.uint8)))
        q.close()
        n_sc += 1
    assert n_sc >= 40
    corpus.close()
