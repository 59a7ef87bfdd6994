#!/usr/bin/env python
"""Writes tests/golden/canonical_vectors.json: vectors produced by THIRD-PARTY code only —
  * base hashes = MurmurHash3_x64_128(key) ++ MurmurHash3_x64_128(key || 0x01) from Austin Appleby's canonical
    MurmurHash3.cpp (oracle/_ref/libmurmur3_canonical.so, compiled unmodified out of scikit-learn's tree), i.e. what
    bloom/v3's sum256 documents itself to equal;
  * CRC32C from the CPU's SSE4.2 crc32 instruction (oracle/_ref/libcrc32c_hw.so).
Neither oracle/bloomref.c nor the CUDA code takes part in generating them.  Run from the repo root:
    python tests/golden/make_canonical_vectors.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import murmur_canonical as canon  # noqa: E402


def main():
    rng = random.Random(20261017)
    keys = [b"", b"a", b"hello", b"service::auth", b"The quick brown fox jumps over the lazy dog", b"\x00" * 16, b"\xff" * 33, b"\x01"]
    for L in list(range(0, 65)) + [79, 80, 95, 96, 127, 128, 129, 255, 256, 257, 1000]:
        keys.append(bytes(rng.randrange(256) for _ in range(L)))
    out = {"_note": "THIRD-PARTY-generated vectors (canonical MurmurHash3.cpp; SSE4.2 crc32 instruction); see make_canonical_vectors.py",
           "base_hashes": [{"key_hex": k.hex(), "h": ["%016x" % x for x in canon.base_hashes(k)]} for k in keys],
           "crc32c": [{"data_hex": k.hex(), "crc": "%08x" % canon.crc32c_hw(k)} for k in keys[:40]]}
    path = os.path.join(ROOT, "tests", "golden", "canonical_vectors.json")
    json.dump(out, open(path, "w"), indent=0)
    print(path, len(out["base_hashes"]), "hash vectors,", len(out["crc32c"]), "crc vectors")


if __name__ == "__main__":
    main()
