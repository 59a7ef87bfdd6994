#!/usr/bin/env python
"""Regenerates tests/golden/bloom_golden.json from the CPU oracle.

The reference (Go, un-vendored bloom/v3) cannot run in the build container, so these
fixtures are NOT outputs of the reference: they freeze the oracle's restatement
(cross-checked C vs Python, murmur3 core pinned by public vectors) so that neither
the oracle nor the CUDA path can drift silently.  "parity unpinned" still applies
at the bit level until a Go run confirms them.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bloomref as py  # noqa: E402
from oracle import cref  # noqa: E402

KEYS = ["", "a", "service", "auth", "service::auth", "user.name", "user.age", "alice", "30",
        "user.name::alice", "user.age::30", "nonexistent.field", "0123456789abcde", "0123456789abcdef",
        "0123456789abcdef0", "timestamp::1700000000", "the quick brown fox jumps over the lazy dog",
        "héllo wörld", "日本語"]


def main():
    out = {"_note": "oracle-generated (NOT reference-generated) fixtures; see make_golden.py"}
    out["base_hashes"] = []
    for k in KEYS:
        kb = k.encode("utf-8")
        h = cref.base_hashes(kb)
        assert h == py.base_hashes(kb)
        out["base_hashes"].append({"key": k, "h": ["%016x" % x for x in h],
                                   "locations_0_11": ["%016x" % cref.location(h, i) for i in range(12)]})
    out["estimate_parameters"] = [{"n": n, "p": p, "m": cref.estimate_parameters(n, p)[0],
                                   "k": cref.estimate_parameters(n, p)[1]} for n, p in
                                  [(1, .001), (2, .001), (2, .02), (100, .01), (101, .001), (1000, .001),
                                   (10 ** 4, .001), (5 * 10 ** 4, .01), (10 ** 5, .001), (10 ** 6, .001), (220, .001),
                                   (9, .001), (1, 1e-9), (12345, 1e-9)]]
    out["filters"] = []
    cases = [
        ("tree_test_fields", 100, 0.01, ["user.name", "user.age"], False),        # bloom_tree_engine_test.go:364-378
        ("tree_test_tokens", 100, 0.01, ["alice", "30"], False),
        ("tree_test_fieldtokens", 100, 0.01, ["user.name::alice", "user.age::30"], False),
        ("sized_two", None, 0.001, ["service", "id"], True),
        ("sized_empty", None, 0.001, [], True),                                     # ingest.go:135-138 empty-set rule
        ("sized_bench_fields", None, 0.001, ["timestamp", "level", "service", "message", "user_id", "nested",
                                             "nested.region", "nested.az", "tags"], True),
    ]
    for name, n, fpr, entries, sized in cases:
        eb = [e.encode() for e in entries]
        f = cref.Filter.build_sized(eb, fpr) if sized else cref.Filter.with_estimates(n, fpr)
        pf = py.build_sized_filter(eb, fpr) if sized else py.BloomFilter.with_estimates(n, fpr)
        if not sized:
            for e in eb:
                f.add(e)
                pf.add(e)
        assert f.write_to() == pf.write_to()
        out["filters"].append({"name": name, "entries": entries, "fpr": fpr, "n": n, "m": f.m, "k": f.k,
                               "words": ["%016x" % int(w) for w in f.words()], "write_to_hex": f.write_to().hex()})
    ff = cref.Filter.build_sized([b"service", b"id"], 0.001)
    tf = cref.Filter.build_sized([b"auth", b"1"], 0.001)
    sec = cref.section_encode(ff, tf, None)
    assert sec == py.encode_filter_section(py.build_sized_filter([b"service", b"id"], 0.001),
                                           py.build_sized_filter([b"auth", b"1"], 0.001), None)
    out["section"] = {"field_entries": ["service", "id"], "token_entries": ["auth", "1"], "fieldtoken": None,
                      "fpr": 0.001, "hex": sec.hex(), "crc32c_check_123456789": "%08x" % cref.crc32c(b"123456789")}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bloom_golden.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1, ensure_ascii=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
