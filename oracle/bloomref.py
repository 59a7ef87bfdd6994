"""Independent pure-Python restatement of the bloom arithmetic — ORACLE, TEST INFRASTRUCTURE ONLY.

Second, deliberately differently-structured statement of what
github.com/bits-and-blooms/bloom/v3 v3.7.0 (+ bitset v1.10.0) computes for the
reference's call sites (ingest.go:139-145 NewWithEstimates/AddString,
query_exec.go:128-159 TestString, file_format.go:368,420 WriteTo/ReadFrom).
It exists to cross-check oracle/bloomref.c: the two must agree bit for bit.

PARITY STATUS: "parity unpinned" at the bit level (see oracle/bloomref.h).  The
murmur3 core here is pinned by the public MurmurHash3_x64_128 vectors in
tests/golden/murmur3_x64_128.json.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
Pure-Python loops: small cases only.
"""
from __future__ import annotations

import math
import struct

MASK64 = (1 << 64) - 1
C1 = 0x87C37B91114253D5
C2 = 0x4CF5AD432745937F


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (64 - r))) & MASK64


def _fmix(k: int) -> int:
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & MASK64
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & MASK64
    k ^= k >> 33
    return k


def murmur3_x64_128(data: bytes, seed: int = 0) -> tuple[int, int]:
    """Austin Appleby's MurmurHash3_x64_128 (public domain), straight from the published algorithm."""
    h1 = h2 = seed & MASK64
    n = len(data)
    nblocks = n // 16
    for b in range(nblocks):
        k1, k2 = struct.unpack_from("<QQ", data, 16 * b)
        k1 = (k1 * C1) & MASK64
        k1 = _rotl(k1, 31)
        k1 = (k1 * C2) & MASK64
        h1 ^= k1
        h1 = _rotl(h1, 27)
        h1 = (h1 + h2) & MASK64
        h1 = (h1 * 5 + 0x52DCE729) & MASK64
        k2 = (k2 * C2) & MASK64
        k2 = _rotl(k2, 33)
        k2 = (k2 * C1) & MASK64
        h2 ^= k2
        h2 = _rotl(h2, 31)
        h2 = (h2 + h1) & MASK64
        h2 = (h2 * 5 + 0x38495AB5) & MASK64
    tail = data[16 * nblocks:]
    if len(tail) > 8:
        k2 = int.from_bytes(tail[8:], "little")
        k2 = (k2 * C2) & MASK64
        k2 = _rotl(k2, 33)
        k2 = (k2 * C1) & MASK64
        h2 ^= k2
    if len(tail) > 0:
        k1 = int.from_bytes(tail[:8], "little")
        k1 = (k1 * C1) & MASK64
        k1 = _rotl(k1, 31)
        k1 = (k1 * C2) & MASK64
        h1 ^= k1
    h1 ^= n
    h2 ^= n
    h1 = (h1 + h2) & MASK64
    h2 = (h2 + h1) & MASK64
    h1 = _fmix(h1)
    h2 = _fmix(h2)
    h1 = (h1 + h2) & MASK64
    h2 = (h2 + h1) & MASK64
    return h1, h2


def base_hashes(data: bytes) -> tuple[int, int, int, int]:
    """bloom.go baseHashes: murmur(data) and murmur(data + b'\\x01'), seed 0.

    (The library computes the second without materialising the byte; its doc comment
    states strict equivalence with hashing data then one more 0x01 byte.)"""
    a = murmur3_x64_128(data)
    b = murmur3_x64_128(data + b"\x01")
    return a[0], a[1], b[0], b[1]


def location(h, i: int) -> int:
    """bloom.go location(): h[i%2] + i*h[2+(((i+(i%2))%4)/2)] in uint64 arithmetic."""
    return (h[i % 2] + i * h[2 + (((i + (i % 2)) % 4) // 2)]) & MASK64


def estimate_parameters(n: int, p: float) -> tuple[int, int]:
    m = int(math.ceil(-1 * float(n) * math.log(p) / math.pow(math.log(2), 2)))
    k = int(math.ceil(math.log(2) * float(m) / float(n)))
    return max(m, 1), max(k, 1)


class BloomFilter:
    """bloom.BloomFilter over a Python int used as the bitset."""

    def __init__(self, m: int, k: int):
        self.m = max(1, m)
        self.k = max(1, k)
        self.bits = 0

    @classmethod
    def with_estimates(cls, n: int, fpr: float) -> "BloomFilter":
        return cls(*estimate_parameters(n, fpr))

    def add(self, data: bytes) -> None:
        h = base_hashes(data)
        for i in range(self.k):
            self.bits |= 1 << (location(h, i) % self.m)

    def test(self, data: bytes) -> bool:
        h = base_hashes(data)
        return all((self.bits >> (location(h, i) % self.m)) & 1 for i in range(self.k))

    @property
    def nwords(self) -> int:
        return (self.m + 63) // 64

    def words(self) -> list[int]:
        return [(self.bits >> (64 * w)) & MASK64 for w in range(self.nwords)]

    def write_to(self) -> bytes:
        """BloomFilter.WriteTo: u64 BE m, u64 BE k, then bitset.WriteTo (u64 BE length, words u64 BE)."""
        return struct.pack(">QQQ", self.m, self.k, self.m) + b"".join(struct.pack(">Q", w) for w in self.words())

    @classmethod
    def read_from(cls, raw: bytes) -> "BloomFilter":
        m, k, bitlen = struct.unpack_from(">QQQ", raw, 0)
        f = cls.__new__(cls)
        f.m, f.k, f.bits = m, k, 0
        nwords = (bitlen + 63) // 64
        for w in range(nwords):
            f.bits |= struct.unpack_from(">Q", raw, 24 + 8 * w)[0] << (64 * w)
        return f


def build_sized_filter(entries, fpr: float) -> BloomFilter:
    """ingest.go:139-145 buildSizedBloomFilter."""
    entries = list(entries)
    f = BloomFilter.with_estimates(max(len(entries), 1), fpr)
    for e in entries:
        f.add(e)
    return f


def crc32c(data: bytes) -> int:
    """Bitwise CRC32C (Castagnoli), file_format.go:44."""
    c = 0xFFFFFFFF
    for b in data:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    return c ^ 0xFFFFFFFF


def encode_filter_section(field, token, fieldtoken) -> bytes:
    """file_format.go:343-385."""
    flags = 0
    body = b""
    for bit, f in enumerate((field, token, fieldtoken)):
        if f is not None:
            flags |= 1 << bit
            raw = f.write_to()
            body += struct.pack("<I", len(raw)) + raw
    payload = bytes([flags]) + body
    return payload + struct.pack("<I", crc32c(payload))


def parse_filter_section(section: bytes):
    """file_format.go:392-448 -> (field, token, fieldtoken) with None for absent."""
    if len(section) < 5:
        raise ValueError("section too small")
    payload, crc = section[:-4], struct.unpack("<I", section[-4:])[0]
    if crc32c(payload) != crc:
        raise ValueError("invalid hash")
    flags = payload[0]
    if flags & ~7:
        raise ValueError("unrecognized flags")
    rest = payload[1:]
    out = []
    for bit in range(3):
        if not flags & (1 << bit):
            out.append(None)
            continue
        if len(rest) < 4:
            raise ValueError("truncated length prefix")
        n = struct.unpack("<I", rest[:4])[0]
        rest = rest[4:]
        if n > len(rest):
            raise ValueError("length exceeds remainder")
        out.append(BloomFilter.read_from(rest[:n]))
        rest = rest[n:]
    if rest:
        raise ValueError("trailing bytes")
    return tuple(out)


# --- expression tree, query_exec.go:75-159 (dict form: see bloomsearch_b200.query) ---
def make_field_token_key(field: bytes, token: bytes) -> bytes:
    """tokenizer.go:508-511."""
    return field + b"::" + token


def evaluate_bloom_filters(field_f, token_f, fieldtoken_f, expr) -> bool:
    """expr: None | ("COND", None) | ("COND", (type, field, token)) | ("AND"|"OR", [children]) | (other, ...)."""
    if expr is None:
        return True
    kind = expr[0]
    if kind == "COND":
        cond = expr[1]
        if cond is None:
            return True
        ctype, field, token = cond
        if ctype == "FIELD":
            return True if field_f is None else field_f.test(field)
        if ctype == "TOKEN":
            return True if token_f is None else token_f.test(token)
        if ctype == "FIELD_TOKEN":
            return True if fieldtoken_f is None else fieldtoken_f.test(make_field_token_key(field, token))
        return False
    if kind == "OR":
        if not expr[1]:
            return False
        return any(evaluate_bloom_filters(field_f, token_f, fieldtoken_f, c) for c in expr[1])
    if kind == "AND":
        return all(evaluate_bloom_filters(field_f, token_f, fieldtoken_f, c) for c in expr[1])
    return False
