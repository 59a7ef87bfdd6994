"""CPU tests pinning the oracle: public known answers, the reference's semantic pins,
committed golden fixtures, and agreement between the two independent restatements."""
from __future__ import annotations

import json
import os
import random

import numpy as np
import pytest

from oracle import bloomref as py
from oracle import cref
from oracle import murmur_canonical as canon

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_murmur3_public_vectors():
    vec = json.load(open(os.path.join(GOLDEN, "murmur3_x64_128.json")))["vectors"]
    assert len(vec) >= 5
    for v in vec:
        want = (int(v["h1"], 16), int(v["h2"], 16))
        data = v["data"].encode()
        assert cref.murmur3_x64_128(data) == want
        assert py.murmur3_x64_128(data) == want
        # baseHashes' first half is exactly murmur(data, seed 0)
        assert cref.base_hashes(data)[:2] == want


def test_base_hashes_second_half_is_murmur_of_data_plus_one():
    # bloom/v3 murmur.go sum256 doc: equivalent to Write(data); Sum128(); Write([]byte{1}); Sum128()
    rng = random.Random(3)
    for L in list(range(0, 50)) + [63, 64, 65, 127, 128, 129, 1000]:
        d = bytes(rng.randrange(256) for _ in range(L))
        h = cref.base_hashes(d)
        assert h[:2] == cref.murmur3_x64_128(d)
        assert h[2:] == cref.murmur3_x64_128(d + b"\x01")
        assert h == py.base_hashes(d)


def _tail_shape_keys(seed):
    """Every length 0..300 (all 16 tail shapes, several block counts), a few long keys, and the
    tail lengths ADVICE r1 singles out (len % 16 in {0, 8, 15}) at several sizes."""
    rng = random.Random(seed)
    lens = list(range(0, 301)) + [1000, 4096, 4097, 70001] + [16 * b + t for b in (7, 40, 200) for t in (0, 8, 15)]
    return [bytes(rng.randrange(256) for _ in range(L)) for L in lens for _ in range(3)] + [b"\x00" * 31, b"\xff" * 47, b"\x01"]


@pytest.mark.skipif(not canon.available(), reason="scikit-learn's copy of MurmurHash3.cpp is absent")
def test_hash_core_matches_canonical_murmurhash3():
    """THIRD-PARTY PIN.  oracle/_ref/libmurmur3_canonical.so is Austin Appleby's MurmurHash3.cpp compiled
    unmodified (oracle/Makefile `ref`); bloom/v3's sum256 documents strict equivalence with
    MurmurHash3_x64_128(data) and MurmurHash3_x64_128(data || 0x01).  Both restatements must agree with it
    for every tail shape, including the virtual 0x01 byte landing at tail offsets 0, 8 and 15."""
    assert canon.murmur3_x64_128(b"hello") == (0xCBD8A7B341BD9B02, 0x5B1E906A48AE1D19)
    vec = json.load(open(os.path.join(GOLDEN, "murmur3_x64_128.json")))["vectors"]
    for v in vec:   # the canonical code reproduces the published known answers
        assert canon.murmur3_x64_128(v["data"].encode()) == (int(v["h1"], 16), int(v["h2"], 16))
    for d in _tail_shape_keys(11):
        want = canon.base_hashes(d)
        assert tuple(cref.base_hashes(d)) == want, len(d)
        if len(d) <= 4097:
            assert tuple(py.base_hashes(d)) == want, len(d)
    for seed in (1, 0x9747B28C, 0xFFFFFFFF):   # the seeded core too (bloom/v3 uses seed 0 only)
        for d in (b"", b"a", b"0123456789abcdef", b"0123456789abcdefg" * 3):
            assert cref.murmur3_x64_128(d, seed) == canon.murmur3_x64_128(d, seed)


@pytest.mark.skipif(not canon.available(), reason="scikit-learn's copy of MurmurHash3.cpp is absent")
def test_fixture_base_hashes_match_canonical_murmurhash3():
    """The committed oracle-pin fixtures (tests/golden/bloom_golden.json) carry the base hashes of 19 keys:
    every one of them equals the canonical code's output, so the fixtures are not merely self-consistent."""
    g = json.load(open(os.path.join(GOLDEN, "bloom_golden.json")))
    assert len(g["base_hashes"]) >= 19
    for e in g["base_hashes"]:
        key = e["key"].encode("utf-8")
        assert tuple(int(x, 16) for x in e["h"]) == canon.base_hashes(key), e["key"]


def test_committed_third_party_vectors():
    """tests/golden/canonical_vectors.json was written by third-party code only (make_canonical_vectors.py: the
    canonical MurmurHash3.cpp and the CPU's crc32 instruction); it needs neither of them at test time, so this
    pin holds on any box.  Both restatements must reproduce every vector."""
    g = json.load(open(os.path.join(GOLDEN, "canonical_vectors.json")))
    assert len(g["base_hashes"]) >= 80 and len(g["crc32c"]) >= 40
    for e in g["base_hashes"]:
        key = bytes.fromhex(e["key_hex"])
        want = tuple(int(x, 16) for x in e["h"])
        assert tuple(cref.base_hashes(key)) == want, e["key_hex"]
        assert tuple(py.base_hashes(key)) == want, e["key_hex"]
    for e in g["crc32c"]:
        assert cref.crc32c_sw(bytes.fromhex(e["data_hex"])) == int(e["crc"], 16)
        assert cref.crc32c(bytes.fromhex(e["data_hex"])) == int(e["crc"], 16)
        assert py.crc32c(bytes.fromhex(e["data_hex"])) == int(e["crc"], 16)


def test_location_pattern():
    # h0; h1+h3; h0+2h3; h1+3h2; h0+4h2; h1+5h3; h0+6h3; h1+7h2
    h = (3, 5, 7, 11)
    want = [3, 5 + 11, 3 + 2 * 11, 5 + 3 * 7, 3 + 4 * 7, 5 + 5 * 11, 3 + 6 * 11, 5 + 7 * 7]
    assert [cref.location(h, i) for i in range(8)] == want
    assert [py.location(h, i) for i in range(8)] == want
    big = (2 ** 64 - 1, 2 ** 64 - 2, 2 ** 63 + 5, 2 ** 64 - 7)
    for i in range(40):
        assert cref.location(big, i) == py.location(big, i) < 2 ** 64


def test_estimate_parameters_table():
    table = [((1, .001), (15, 11)), ((2, .001), (29, 11)), ((2, .02), (17, 6)), ((100, .01), (959, 7)),
             ((101, .001), (1453, 10)), ((1000, .001), (14378, 10)), ((10 ** 4, .001), (143776, 10)),
             ((5 * 10 ** 4, .01), (479253, 7)), ((10 ** 5, .001), (1437759, 10)), ((10 ** 6, .001), (14377588, 10))]
    for (n, p), want in table:
        assert cref.estimate_parameters(n, p) == want
        assert py.estimate_parameters(n, p) == want


@pytest.mark.skipif(not canon.crc32c_hw_available(), reason="no SSE4.2 crc32 instruction / gcc on this host")
def test_crc32c_matches_the_cpu_instruction():
    """THIRD-PARTY PIN: the oracle's table-driven CRC32C (bref_crc32c_sw: the restated algorithm; the checksum
    of every filter section, file_format.go:44,379,399: crc32.Checksum(payload, Castagnoli)) and the Python
    twin's against the CPU's SSE4.2 crc32 instruction, for every length 0..300, long buffers and the standard
    check value.  (bref_crc32c itself takes the instruction where the host has it, as Go's hash/crc32 does — so
    the comparison that means something is the software path's.)"""
    assert canon.crc32c_hw(b"123456789") == 0xE3069283 == cref.crc32c_sw(b"123456789")
    rng = random.Random(5)
    for L in list(range(0, 301)) + [1000, 4096, 65537, 1 << 20]:
        d = bytes(rng.getrandbits(8) for _ in range(L)) if L <= 65537 else rng.randbytes(L)
        want = canon.crc32c_hw(d)
        assert cref.crc32c_sw(d) == want, L
        assert cref.crc32c(d) == want, L
        if L <= 4096:
            assert py.crc32c(d) == want, L


def test_crc32c_check_value():
    assert cref.crc32c(b"123456789") == 0xE3069283 == py.crc32c(b"123456789")
    rng = random.Random(1)
    for L in (0, 1, 7, 8, 9, 63, 64, 1000):
        d = bytes(rng.randrange(256) for _ in range(L))
        assert cref.crc32c(d) == py.crc32c(d)


def test_oracle_pin_fixtures_still_hold():
    g = json.load(open(os.path.join(GOLDEN, "bloom_golden.json")))
    for e in g["base_hashes"]:
        kb = e["key"].encode("utf-8")
        want = tuple(int(x, 16) for x in e["h"])
        assert cref.base_hashes(kb) == want == py.base_hashes(kb)
        assert [cref.location(want, i) for i in range(12)] == [int(x, 16) for x in e["locations_0_11"]]
    for e in g["estimate_parameters"]:
        assert cref.estimate_parameters(e["n"], e["p"]) == (e["m"], e["k"])
    for e in g["filters"]:
        eb = [x.encode() for x in e["entries"]]
        if e["n"] is None:
            f = cref.Filter.build_sized(eb, e["fpr"])
        else:
            f = cref.Filter.with_estimates(e["n"], e["fpr"])
            for x in eb:
                f.add(x)
        assert (f.m, f.k) == (e["m"], e["k"])
        assert ["%016x" % int(w) for w in f.words()] == e["words"]
        assert f.write_to().hex() == e["write_to_hex"]
    s = g["section"]
    ff = cref.Filter.build_sized([x.encode() for x in s["field_entries"]], s["fpr"])
    tf = cref.Filter.build_sized([x.encode() for x in s["token_entries"]], s["fpr"])
    assert cref.section_encode(ff, tf, None).hex() == s["hex"]


def test_filter_sizing_and_empty_set_rule():
    # ingest.go:135-145: empty set is sized for one entry and tests negative
    f = cref.Filter.build_sized([], 0.001)
    assert (f.m, f.k) == (15, 11) and not f.words().any()
    assert not f.test(b"anything")
    # file_format_test.go:28-94 counts {2,101,101}
    for n in (2, 101):
        keys = [b"k%d" % i for i in range(n)]
        f = cref.Filter.build_sized(keys, 0.001)
        assert (f.m, f.k) == cref.estimate_parameters(n, 0.001)
        assert all(f.test(k) for k in keys)


def test_c_and_python_filters_agree():
    rng = random.Random(5)
    for n in (1, 2, 10, 100, 500):
        keys = sorted({bytes(rng.randrange(256) for _ in range(rng.randint(0, 30))) for _ in range(n)})
        cf = cref.Filter.build_sized(keys, 0.01)
        pf = py.build_sized_filter(keys, 0.01)
        assert (cf.m, cf.k) == (pf.m, pf.k)
        assert [int(w) for w in cf.words()] == pf.words()
        assert cf.write_to() == pf.write_to()
        probe = keys[:5] + [b"absent-%d" % i for i in range(20)]
        assert [cf.test(k) for k in probe] == [pf.test(k) for k in probe]


def test_measured_fpr_within_budget():
    # file_format_test.go:100-165: 50 000 distinct tokens, fpr 0.01, 10 000 absent probes, <= 3x
    keys = [b"token-%d" % i for i in range(50_000)]
    f = cref.Filter.build_sized(keys, 0.01)
    assert (f.m, f.k) == (479253, 7)
    blob, off = cref.pack_keys([b"missing-%d" % i for i in range(10_000)])
    desc = np.array([(0, 0, 0), (f.m, f.k, 0), (0, 0, 0)], dtype=cref.DESC_DTYPE)
    m = cref.probe_matrix(desc, f.words(), 1, blob, off, np.ones(10_000, np.uint8))
    fp = int(np.unpackbits(m.view(np.uint8)).sum())
    assert fp <= 3 * 0.01 * 10_000
    blob, off = cref.pack_keys(keys[:2000])
    m = cref.probe_matrix(desc, f.words(), 1, blob, off, np.ones(2000, np.uint8))
    assert int(np.unpackbits(m.view(np.uint8)).sum()) == 2000  # no false negatives


def test_write_read_round_trip_and_section_codec():
    # file_format_test.go:439-443 (filters Equal after round trip) + framing errors :583-802
    ff = cref.Filter.build_sized([b"a", b"b"], 0.001)
    tf = cref.Filter.build_sized([b"x%d" % i for i in range(300)], 0.001)
    assert cref.Filter.read_from(tf.write_to()).equal(tf)
    assert py.BloomFilter.read_from(tf.write_to()).write_to() == tf.write_to()
    for combo in [(ff, tf, None), (None, None, None), (None, tf, ff), (ff, None, None)]:
        sec = cref.section_encode(*combo)
        pcombo = tuple(None if f is None else py.BloomFilter.read_from(f.write_to()) for f in combo)
        assert sec == py.encode_filter_section(*pcombo)
        back = cref.section_parse(sec)
        for a, b in zip(combo, back):
            assert (a is None) == (b is None)
            if a is not None:
                assert a.equal(b)
        pback = py.parse_filter_section(sec)
        for a, b in zip(combo, pback):
            assert (a is None) == (b is None)
    sec = bytearray(cref.section_encode(ff, tf, None))
    sec[10] ^= 0x40  # byte flip -> CRC mismatch
    with pytest.raises(ValueError):
        cref.section_parse(bytes(sec))
    with pytest.raises(ValueError):
        py.parse_filter_section(bytes(sec))
    with pytest.raises(ValueError):
        cref.section_parse(b"\x00\x00")  # too small


def _tree_cases():
    F, T, FT = (lambda f: ("COND", ("FIELD", f, b""))), (lambda t: ("COND", ("TOKEN", b"", t))), \
        (lambda f, t: ("COND", ("FIELD_TOKEN", f, t)))
    return [
        (None, True),
        (F(b"user.name"), True),
        (F(b"nonexistent.field"), False),
        (T(b"alice"), True),
        (FT(b"user.name", b"alice"), True),
        (("OR", [F(b"nonexistent.field"), F(b"user.name")]), True),
        (("AND", [F(b"nonexistent.field"), F(b"user.name")]), False),
        (("OR", [F(b"nonexistent.field"), FT(b"user.name", b"alice")]), True),
        # semantics beyond the 8 reference cases (query_exec.go:97-125,155-156)
        (("OR", []), False),
        (("AND", []), True),
        (("COND", None), True),
        (("BOGUS", []), False),
        (("COND", ("BOGUS", b"x", b"y")), False),
        (("AND", [("OR", [T(b"nope"), T(b"30")]), ("AND", []), FT(b"user.age", b"30")]), True),
    ]


def test_evaluate_bloom_filters_reference_semantics():
    # bloom_tree_engine_test.go:357-442 with the same filters (NewWithEstimates(100, 0.01), 2 entries each)
    def mk(cls, entries):
        f = cls.with_estimates(100, 0.01)
        for e in entries:
            f.add(e)
        return f
    sets = ([b"user.name", b"user.age"], [b"alice", b"30"], [b"user.name::alice", b"user.age::30"])
    cf = [mk(cref.Filter, s) for s in sets]
    pf = [mk(py.BloomFilter, s) for s in sets]
    for expr, want in _tree_cases():
        assert cref.evaluate_bloom_filters(*cf, expr) == want, expr
        assert py.evaluate_bloom_filters(*pf, expr) == want, expr
    # nil filters cannot disqualify (query_exec.go:137-151)
    assert cref.evaluate_bloom_filters(None, cf[1], cf[2], ("COND", ("FIELD", b"nonexistent.field", b""))) is True
    assert py.evaluate_bloom_filters(None, None, None, ("AND", [("COND", ("TOKEN", b"", b"zzz"))])) is True


def test_probe_sections_matches_probe_matrix():
    rng = random.Random(8)
    units = []
    for u in range(12):
        units.append(tuple(sorted({bytes(rng.randrange(97, 123) for _ in range(rng.randint(1, 9)))
                                   for _ in range(40 + u)}) for _ in range(3)))
    from tests.helpers import oracle_units
    desc, words = oracle_units(units, 0.01, absent={(2, 1), (5, 0)})
    keys = [units[0][1][0], units[3][2][1], units[7][0][2], b"nope", b"zzzz"]
    kinds = np.array([1, 2, 0, 1, 2], dtype=np.uint8)
    blob, off = cref.pack_keys(keys)
    want = cref.probe_matrix(desc, words, len(units), blob, off, kinds)
    sec, sec_off = cref.encode_sections(desc, words, len(units))
    got, errs = cref.probe_sections_matrix(sec, sec_off, blob, off, kinds, n_threads=3)
    assert errs == 0 and np.array_equal(got, want)
    expr = ("OR", [("COND", ("TOKEN", b"", keys[0])), ("COND", ("FIELD", keys[2], b""))])
    mask, errs = cref.probe_sections(sec, sec_off, expr, n_threads=2)
    prog = np.array([(0, 0), (0, 2), (2, 2)], dtype=cref.OP_DTYPE)
    want_mask = cref.probe_mask(desc, words, len(units), blob, off, kinds, prog)
    assert errs == 0 and np.array_equal(mask, want_mask)
