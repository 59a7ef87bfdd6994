"""Host-side mirror of the reference's filter build / probe call sites over the C ABI.

  BloomEntrySets.build_filters  <- ingest.go:24-145 (bloomEntrySets, buildFilters,
                                   buildSizedBloomFilter), called per block and per file
                                   from flush.go:204,253 / merge.go:516,771
  Corpus.evaluate_bloom_filters <- query_exec.go:75-159, batched over every unit of a
                                   resident corpus instead of one filter triple at a time
  BloomFilter / BloomFilters    <- the value types of file_format.go:328-332 (m, k, words)

All arithmetic happens in libbloomgpu.so on the GPU; nothing here hashes or tests
bits on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _native as N
from .query import BloomQuery, CompiledQuery, compile_bloom_query


class Context:
    """Owns a bsg_ctx (device, streams)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        N.check(N.lib().bsg_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            N.lib().bsg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def device_info(self) -> dict:
        sm, cmaj, cmin = C.c_int(), C.c_int(), C.c_int()
        smem, total = C.c_size_t(), C.c_size_t()
        N.check(N.lib().bsg_device_info(self._h, C.byref(sm), C.byref(smem), C.byref(total), C.byref(cmaj), C.byref(cmin)))
        return {"sm_count": sm.value, "smem_optin": smem.value, "total_mem": total.value,
                "cc": (cmaj.value, cmin.value)}

    def set_stream(self, cuda_stream: Optional[int]):
        N.check(N.lib().bsg_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def synchronize(self):
        N.check(N.lib().bsg_synchronize(self._h))

    def timer_begin(self):
        N.check(N.lib().bsg_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_float()
        N.check(N.lib().bsg_timer_end(self._h, C.byref(ms)))
        return ms.value

    # ---- K1 ----
    def hash_keys(self, keys: Sequence[bytes]) -> np.ndarray:
        blob, off = N.pack_keys(list(keys))
        out = np.zeros((len(keys), 4), dtype=np.uint64)
        N.check(N.lib().bsg_hash_keys(self._h, N.ptr(blob), N.ptr(off), len(keys), N.ptr(out)))
        return out

    # ---- K2/K3 ----
    def build(self, blob: np.ndarray, key_off: np.ndarray, group_begin: np.ndarray, group_filter: np.ndarray,
              group_filter2: Optional[np.ndarray], desc: np.ndarray, n_words: int) -> np.ndarray:
        """Raw bsg_build over packed keys; returns the native-endian words array."""
        desc = np.ascontiguousarray(desc, dtype=N.DESC_DTYPE)
        group_begin = np.ascontiguousarray(group_begin, dtype=np.uint64)
        group_filter = np.ascontiguousarray(group_filter, dtype=np.uint32)
        gf2 = None if group_filter2 is None else np.ascontiguousarray(group_filter2, dtype=np.uint32)
        key_off = np.ascontiguousarray(key_off, dtype=np.uint64)
        out = np.zeros(max(int(n_words), 1), dtype=np.uint64)
        N.check(N.lib().bsg_build(self._h, N.ptr(blob), N.ptr(key_off), len(key_off) - 1, N.ptr(group_begin),
                                  len(group_filter), N.ptr(group_filter), N.ptr(gf2), N.ptr(desc), len(desc),
                                  N.ptr(out), int(n_words)))
        return out[:int(n_words)]

    def build_fieldtokens(self, strings_blob: np.ndarray, str_off: np.ndarray, pair_path: np.ndarray,
                          pair_token: np.ndarray, group_begin: np.ndarray, group_filter: np.ndarray,
                          group_filter2: Optional[np.ndarray], desc: np.ndarray, n_words: int) -> np.ndarray:
        """bsg_build_fieldtokens: entries are (path, token) index pairs into one string table; the
        key path + "::" + token is hashed on the device without being materialised."""
        desc = np.ascontiguousarray(desc, dtype=N.DESC_DTYPE)
        str_off = np.ascontiguousarray(str_off, dtype=np.uint64)
        pair_path = np.ascontiguousarray(pair_path, dtype=np.uint32)
        pair_token = np.ascontiguousarray(pair_token, dtype=np.uint32)
        group_begin = np.ascontiguousarray(group_begin, dtype=np.uint64)
        group_filter = np.ascontiguousarray(group_filter, dtype=np.uint32)
        gf2 = None if group_filter2 is None else np.ascontiguousarray(group_filter2, dtype=np.uint32)
        out = np.zeros(max(int(n_words), 1), dtype=np.uint64)
        N.check(N.lib().bsg_build_fieldtokens(self._h, N.ptr(strings_blob), N.ptr(str_off), len(str_off) - 1,
                                              N.ptr(pair_path), N.ptr(pair_token), len(pair_path), N.ptr(group_begin),
                                              len(group_filter), N.ptr(group_filter), N.ptr(gf2), N.ptr(desc), len(desc),
                                              N.ptr(out), int(n_words)))
        return out[:int(n_words)]

    def count_distinct(self, blob: np.ndarray, key_off: np.ndarray, group_begin: np.ndarray,
                       group_parent: Optional[np.ndarray] = None, n_parents: int = 0):
        """bsg_count_distinct: exact distinct-entry counts per group (and per parent union) of emissions
        that may repeat — the counts bloomEntrySets' maps provide (ingest.go:24-45,105-123)."""
        key_off = np.ascontiguousarray(key_off, dtype=np.uint64)
        group_begin = np.ascontiguousarray(group_begin, dtype=np.uint64)
        n_groups = len(group_begin) - 1
        gp = None if group_parent is None else np.ascontiguousarray(group_parent, dtype=np.uint32)
        gc = np.zeros(max(n_groups, 1), dtype=np.uint64)
        pc = None if gp is None else np.zeros(max(int(n_parents), 1), dtype=np.uint64)
        N.check(N.lib().bsg_count_distinct(self._h, N.ptr(blob), N.ptr(key_off), len(key_off) - 1, N.ptr(group_begin),
                                           n_groups, N.ptr(gp), int(n_parents), N.ptr(gc), N.ptr(pc)))
        return gc[:n_groups], (None if pc is None else pc[:int(n_parents)])

    # ---- multi-GPU ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        N.check(N.lib().bsg_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        N.check(N.lib().bsg_comm_init(self._h, rank, world, buf))

    def or_reduce(self, words: np.ndarray) -> np.ndarray:
        words = np.ascontiguousarray(words, dtype=np.uint64)
        N.check(N.lib().bsg_or_reduce(self._h, N.ptr(words), len(words)))
        return words

    def allgather_masks(self, local: np.ndarray, world: int) -> np.ndarray:
        local = np.ascontiguousarray(local, dtype=np.uint64)
        out = np.zeros((world, len(local)), dtype=np.uint64)
        N.check(N.lib().bsg_allgather_masks(self._h, N.ptr(local), len(local), N.ptr(out)))
        return out

    def read_device(self, dev_ptr: int, n_words: int) -> np.ndarray:
        """Synchronise the ctx stream and copy n_words uint64 of device memory to the host (tests / bench)."""
        out = np.zeros(max(int(n_words), 1), dtype=np.uint64)
        L = N.lib()
        L.bsg_debug_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        N.check(L.bsg_debug_memcpy_d2h(self._h, N.ptr(out), C.c_void_p(dev_ptr), int(n_words) * 8))
        return out[:int(n_words)]

    def host_alloc(self, shape, dtype=np.uint64) -> np.ndarray:
        """A numpy array over pinned, device-mapped host memory (bsg_host_alloc): pass it as out_matrix and the
        probe kernel writes the rows straight into it.  Free with host_free(array)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        N.check(N.lib().bsg_host_alloc(self._h, n, C.byref(p)))
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        arr.flags.writeable = True
        self._host_ptrs = getattr(self, "_host_ptrs", {})
        self._host_ptrs[arr.ctypes.data] = p.value
        return arr

    def host_free(self, arr: np.ndarray):
        p = getattr(self, "_host_ptrs", {}).pop(arr.ctypes.data, None)
        if p is not None:
            N.check(N.lib().bsg_host_free(self._h, C.c_void_p(p)))

    def comm_info(self) -> dict:
        r, w, p, b = C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
        N.check(N.lib().bsg_comm_info(self._h, C.byref(r), C.byref(w), C.byref(p), C.byref(b)))
        return {"rank": r.value, "world": w.value, "peer_memory": bool(p.value), "last_nvlink_bytes": b.value}

    def comm_alloc(self, nbytes: int) -> int:
        """Symmetric device memory (collective); returns this rank's device pointer."""
        p = C.c_void_p()
        N.check(N.lib().bsg_comm_alloc(self._h, int(nbytes), C.byref(p)))
        return p.value

    def comm_free(self, dev_ptr: int):
        N.check(N.lib().bsg_comm_free(self._h, C.c_void_p(dev_ptr)))

    def or_reduce_device(self, dev_ptr: int, n_words: int):
        """In-place OR across ranks of a symmetric device buffer; asynchronous on the ctx stream."""
        N.check(N.lib().bsg_or_reduce_device(self._h, C.c_void_p(dev_ptr), int(n_words)))

    def allgather_masks_device(self, d_local: int, n_words: int, d_all: int):
        N.check(N.lib().bsg_allgather_masks_device(self._h, C.c_void_p(d_local), int(n_words), C.c_void_p(d_all)))


def estimate_parameters(n: int, fpr: float) -> tuple[int, int]:
    """bloom.EstimateParameters + New's clamp (bsg_estimate)."""
    m, k = C.c_uint64(), C.c_uint64()
    N.lib().bsg_estimate(n, fpr, C.byref(m), C.byref(k))
    return m.value, k.value


@dataclass
class BloomFilter:
    """Value mirror of *bloom.BloomFilter: m bits, k hashes, native-endian uint64 words."""
    m: int
    k: int
    words: np.ndarray

    def write_to(self) -> bytes:
        """BloomFilter.WriteTo framing (big-endian), file_format.go:368."""
        hdr = np.array([self.m, self.k, self.m], dtype=">u8").tobytes()
        return hdr + self.words.astype(">u8").tobytes()

    @classmethod
    def read_from(cls, raw: bytes) -> "BloomFilter":
        m, k, bitlen = (int(x) for x in np.frombuffer(raw[:24], dtype=">u8"))
        nw = (bitlen + 63) // 64
        return cls(m, k, np.frombuffer(raw[24:24 + 8 * nw], dtype=">u8").astype(np.uint64))


@dataclass
class BloomFilters:
    """file_format.go:328-332; any member may be None (absent => cannot disqualify)."""
    FieldBloomFilter: Optional[BloomFilter] = None
    TokenBloomFilter: Optional[BloomFilter] = None
    FieldTokenBloomFilter: Optional[BloomFilter] = None

    def as_tuple(self):
        return (self.FieldBloomFilter, self.TokenBloomFilter, self.FieldTokenBloomFilter)


class BloomEntrySets:
    """ingest.go:24-123: distinct field / token / field::token entries of a set of rows.
    Dedup stays on the host (Python sets here, Go maps in the reference); hashing and
    bit-setting run on the GPU at build_filters time."""

    def __init__(self):
        self.fields: set = set()
        self.tokens: set = set()
        self.fieldTokens: set = set()

    def add_field(self, path: bytes):
        self.fields.add(path)

    def add_token(self, token: bytes):
        self.tokens.add(token)

    def add_field_token(self, path: bytes, token: bytes):  # ingest.go:95-102
        self.fieldTokens.add(path + b"::" + token)

    def union_into(self, dst: "BloomEntrySets"):  # ingest.go:105-115
        dst.fields |= self.fields
        dst.tokens |= self.tokens
        dst.fieldTokens |= self.fieldTokens

    def counts(self):  # ingest.go:117-123
        return {"Fields": len(self.fields), "Tokens": len(self.tokens), "FieldTokens": len(self.fieldTokens)}

    def build_filters(self, ctx: Context, false_positive_rate: float) -> BloomFilters:  # ingest.go:127-133
        return build_filters_many(ctx, [self], false_positive_rate)[0]


def build_filters_many(ctx: Context, sets: Sequence[BloomEntrySets], fpr: float,
                       file_sets: Optional[Sequence[BloomEntrySets]] = None,
                       file_of: Optional[Sequence[int]] = None):
    """One bsg_build call for many entry sets (all partition buffers of a flush,
    flush.go:138-282).  If file_sets/file_of are given, set i also feeds the union filter
    of file file_of[i], sized from file_sets[...]'s exact distinct counts (flush.go:221,253);
    returns (block_filters, file_filters) then, else block_filters."""
    keys: List[bytes] = []
    group_begin = [0]
    group_filter: List[int] = []
    group_filter2: List[int] = []
    desc = []
    word_off = 0

    def new_filter(n_entries: int) -> int:
        nonlocal word_off
        m, k = estimate_parameters(max(n_entries, 1), fpr)  # ingest.go:139-140
        desc.append((m, k, word_off))
        word_off += (m + 63) // 64
        return len(desc) - 1

    file_ids = None
    if file_sets is not None:
        file_ids = [[new_filter(len(s)) for s in (fs.fields, fs.tokens, fs.fieldTokens)] for fs in file_sets]
    block_ids = []
    for i, es in enumerate(sets):
        ids = []
        for kind, entries in enumerate((es.fields, es.tokens, es.fieldTokens)):
            fid = new_filter(len(entries))
            ids.append(fid)
            keys.extend(entries)
            group_begin.append(len(keys))
            group_filter.append(fid)
            group_filter2.append(file_ids[file_of[i]][kind] if file_ids is not None else N.NO_FILTER)
        block_ids.append(ids)
    blob, off = N.pack_keys(keys)
    d = np.array(desc, dtype=N.DESC_DTYPE) if desc else np.zeros(0, N.DESC_DTYPE)
    words = ctx.build(blob, off, np.array(group_begin, np.uint64), np.array(group_filter, np.uint32),
                      np.array(group_filter2, np.uint32) if file_ids is not None else None, d, word_off)

    def mk(fid):
        m, k, wo = desc[fid]
        return BloomFilter(m, k, words[wo:wo + (m + 63) // 64].copy())

    blocks = [BloomFilters(*(mk(f) for f in ids)) for ids in block_ids]
    if file_ids is None:
        return blocks
    return blocks, [BloomFilters(*(mk(f) for f in ids)) for ids in file_ids]


class Corpus:
    """A set of units (data blocks or files) whose filters are resident in HBM."""

    def __init__(self, ctx: Context, desc: np.ndarray, words: np.ndarray, big_endian: bool = False):
        desc = np.ascontiguousarray(desc, dtype=N.DESC_DTYPE).reshape(-1)
        assert len(desc) % 3 == 0
        words = np.ascontiguousarray(words, dtype=np.uint64)
        self.ctx = ctx
        self.n_units = len(desc) // 3
        self._h = C.c_void_p()
        N.check(N.lib().bsg_corpus_load(ctx.handle, N.ptr(desc), self.n_units, N.ptr(words) if len(words) else None,
                                        len(words), 1 if big_endian else 0, C.byref(self._h)))

    @classmethod
    def from_sections(cls, ctx: Context, sections: np.ndarray, sec_off: np.ndarray, verify_crc: bool = True):
        """Load straight from raw on-disk filter sections (file_format.go:343-385 framing): framing
        parse, CRC32C check and big-endian decode all run on the device.  Returns (corpus, status)
        where status[u] is 0 or the negative reason unit u failed to parse; failed units keep no
        filters and therefore can never be disqualified (query_exec.go:580-590 error isolation)."""
        sections = np.ascontiguousarray(sections, dtype=np.uint8)
        sec_off = np.ascontiguousarray(sec_off, dtype=np.uint64)
        self = cls.__new__(cls)
        self.ctx = ctx
        self.n_units = len(sec_off) - 1
        self._h = C.c_void_p()
        status = np.zeros(max(self.n_units, 1), dtype=np.int32)
        n_bad = C.c_uint64()
        N.check(N.lib().bsg_corpus_load_sections(ctx.handle, N.ptr(sections) if len(sections) else None, N.ptr(sec_off),
                                                 self.n_units, 1 if verify_crc else 0, N.ptr(status), C.byref(n_bad),
                                                 C.byref(self._h)))
        self.n_bad = n_bad.value
        return self, status[:self.n_units]

    def unit_desc(self, unit: int) -> np.ndarray:
        out = np.zeros(3, dtype=N.DESC_DTYPE)
        N.check(N.lib().bsg_corpus_unit_desc(self._h, unit, N.ptr(out)))
        return out

    @classmethod
    def from_filters(cls, ctx: Context, units: Sequence[BloomFilters]) -> "Corpus":
        desc = np.zeros(len(units) * 3, dtype=N.DESC_DTYPE)
        chunks = []
        off = 0
        for u, bf in enumerate(units):
            for kind, f in enumerate(bf.as_tuple()):
                if f is None:
                    continue
                desc[u * 3 + kind] = (f.m, f.k, off)
                chunks.append(np.asarray(f.words, dtype=np.uint64))
                off += len(f.words)
        words = np.concatenate(chunks) if chunks else np.zeros(0, np.uint64)
        return cls(ctx, desc, words)

    def close(self):
        if self._h and not getattr(self, "_borrowed", False):   # a cache-owned corpus is released, not freed
            N.lib().bsg_corpus_free(self._h)
        self._h = C.c_void_p()

    def device_bytes(self) -> int:
        return N.lib().bsg_corpus_device_bytes(self._h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def bitset_bytes(self, kind_mask: int = 7) -> int:
        return N.lib().bsg_corpus_bitset_bytes(self._h, kind_mask)

    # ---- probe: host buffers in, host buffers out (the call the Go shim makes) ----
    def probe(self, keys: Sequence[bytes], kinds, prog: Optional[np.ndarray] = None, want_matrix: bool = True,
              want_mask: bool = True):
        blob, off = N.pack_keys(list(keys))
        kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
        q = len(keys)
        matrix = np.zeros((self.n_units, (q + 63) // 64), dtype=np.uint64) if want_matrix else None
        mask = np.zeros((self.n_units + 63) // 64, dtype=np.uint64) if want_mask else None
        pp, pl = (None, 0) if prog is None else (N.ptr(np.ascontiguousarray(prog, N.OP_DTYPE)), len(prog))
        N.check(N.lib().bsg_probe(self.ctx.handle, self._h, N.ptr(blob), N.ptr(off), q, N.ptr(kinds), pp, pl,
                                  N.ptr(matrix), N.ptr(mask)))
        return matrix, mask

    def probe_packed(self, blob: np.ndarray, off: np.ndarray, kinds: np.ndarray, prog: Optional[np.ndarray],
                     out_matrix: Optional[np.ndarray], out_mask: Optional[np.ndarray]) -> None:
        """bsg_probe on caller-packed host buffers, results into caller-allocated host arrays
        (exactly the C-ABI call, no Python-side packing)."""
        pp, pl = (None, 0) if prog is None else (N.ptr(prog), len(prog))
        N.check(N.lib().bsg_probe(self.ctx.handle, self._h, N.ptr(blob), N.ptr(off), len(off) - 1, N.ptr(kinds), pp, pl,
                                  N.ptr(out_matrix), N.ptr(out_mask)))

    def probe_multi(self, queries: Sequence[Optional[BloomQuery]]) -> np.ndarray:
        """bsg_probe_multi: several BloomQueries in one pass over the corpus -> bool[n_queries][n_units], row j
        identical to evaluate_bloom_filters(queries[j])."""
        cqs = [compile_bloom_query(q) for q in queries]
        blob, off, kinds, qbegin, progs, pbegin = pack_queries(cqs)
        words = (self.n_units + 63) // 64
        out = np.zeros((len(cqs), max(words, 1)), dtype=np.uint64)
        N.check(N.lib().bsg_probe_multi(self.ctx.handle, self._h, N.ptr(blob), N.ptr(off), len(off) - 1, N.ptr(kinds),
                                        len(cqs), N.ptr(qbegin), N.ptr(progs), N.ptr(pbegin), N.ptr(out)))
        return np.stack([unpack_mask(out[j], self.n_units) for j in range(len(cqs))]) if cqs else np.zeros((0, self.n_units), bool)

    def set_parents(self, parent: np.ndarray, n_parent_units: int) -> None:
        """Record, for every unit (block), the index of its parent unit (file) in another corpus."""
        parent = np.ascontiguousarray(parent, dtype=np.uint32)
        N.check(N.lib().bsg_corpus_set_parents(self.ctx.handle, self._h, N.ptr(parent), len(parent), n_parent_units))

    def evaluate_bloom_filters(self, query: Optional[BloomQuery]) -> np.ndarray:
        """evaluateBloomFilters (query_exec.go:75-87) for every unit at once -> bool[n_units]."""
        cq = compile_bloom_query(query)
        _, mask = self.probe(cq.keys, cq.kinds, cq.prog, want_matrix=False)
        return unpack_mask(mask, self.n_units)


def pack_queries(cqs):
    """CompiledQuery list -> the packed arrays of bsg_probe_multi (keys of all queries back to back, CSR of keys and
    of postfix ops per query; leaf arguments stay query-local)."""
    keys, kinds, qbegin, pbegin, ops = [], [], [0], [0], []
    for cq in cqs:
        keys.extend(cq.keys)
        kinds.extend(int(k) for k in cq.kinds)
        qbegin.append(len(keys))
        if cq.prog is not None:
            ops.append(np.ascontiguousarray(cq.prog, N.OP_DTYPE))
        pbegin.append(pbegin[-1] + (0 if cq.prog is None else len(cq.prog)))
    blob, off = N.pack_keys(keys)
    progs = np.concatenate(ops) if ops else np.zeros(1, N.OP_DTYPE)
    return (blob, off, np.array(kinds if kinds else [0], dtype=np.uint8), np.array(qbegin, dtype=np.uint32),
            np.ascontiguousarray(progs, N.OP_DTYPE), np.array(pbegin, dtype=np.uint32))


class Batcher:
    """bsg_batcher: merges concurrent evaluate() callers on one corpus into bsg_probe_multi launches (group commit).
    Thread-safe; ctypes releases the GIL for the duration of the call, so Python threads really overlap."""

    def __init__(self, corpus: "Corpus", max_keys: int = 0, max_queries: int = 0, window_us: int = 0):
        self.corpus = corpus
        h = C.c_void_p()
        N.check(N.lib().bsg_batcher_create(corpus.ctx.handle, corpus.handle, max_keys, max_queries, window_us, C.byref(h)))
        self._h = h

    def evaluate_packed(self, blob, off, kinds, prog, out_mask) -> None:
        pp, pl = (None, 0) if prog is None else (N.ptr(prog), len(prog))
        N.check(N.lib().bsg_batcher_probe(self._h, N.ptr(blob), N.ptr(off), len(off) - 1, N.ptr(kinds), pp, pl,
                                          N.ptr(out_mask)))

    def evaluate(self, query: Optional[BloomQuery]) -> np.ndarray:
        """evaluateBloomFilters (query_exec.go:75-87) for every unit -> bool[n_units]; may share its launch with
        other threads' queries."""
        cq = compile_bloom_query(query)
        blob, off = N.pack_keys(list(cq.keys))
        kinds = np.ascontiguousarray(cq.kinds if len(cq.keys) else [0], dtype=np.uint8)
        prog = None if cq.prog is None else np.ascontiguousarray(cq.prog, N.OP_DTYPE)
        mask = np.zeros(max((self.corpus.n_units + 63) // 64, 1), dtype=np.uint64)
        self.evaluate_packed(blob, off, kinds, prog, mask)
        return unpack_mask(mask, self.corpus.n_units)

    def stats(self) -> dict:
        v = [C.c_uint64() for _ in range(4)]
        N.check(N.lib().bsg_batcher_stats(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("calls", "launches", "bypassed", "largest_batch"), (int(x.value) for x in v)))

    def close(self):
        if self._h:
            N.lib().bsg_batcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FilterCache:
    """Resident filter cache (bsg_cache): corpora keyed by file id under a byte budget, LRU eviction, pinned
    while in use, invalidated when a merge replaces the file or it is tombstoned (merge.go:529-536)."""

    def __init__(self, ctx: Context, budget_bytes: int):
        self.ctx = ctx
        self._h = C.c_void_p()
        N.check(N.lib().bsg_cache_create(ctx.handle, int(budget_bytes), C.byref(self._h)))

    def _borrow(self, handle) -> "Corpus":
        cp = Corpus.__new__(Corpus)
        cp.ctx = self.ctx
        cp._h = C.c_void_p(handle)
        cp.n_units = N.lib().bsg_corpus_units(cp._h)
        cp._borrowed = True
        return cp

    def acquire(self, file_id: int) -> Optional["Corpus"]:
        out = C.c_void_p()
        N.check(N.lib().bsg_cache_acquire(self._h, int(file_id), C.byref(out)))
        return self._borrow(out.value) if out.value else None

    def insert_sections(self, file_id: int, sections: np.ndarray, sec_off: np.ndarray, verify_crc: bool = True):
        sections = np.ascontiguousarray(sections, dtype=np.uint8)
        sec_off = np.ascontiguousarray(sec_off, dtype=np.uint64)
        n_units = len(sec_off) - 1
        status = np.zeros(max(n_units, 1), dtype=np.int32)
        n_bad, out = C.c_uint64(), C.c_void_p()
        N.check(N.lib().bsg_cache_insert_sections(self._h, int(file_id), N.ptr(sections) if len(sections) else None,
                                                  N.ptr(sec_off), n_units, 1 if verify_crc else 0, N.ptr(status),
                                                  C.byref(n_bad), C.byref(out)))
        return self._borrow(out.value), status[:n_units]

    def insert(self, file_id: int, corpus: "Corpus") -> "Corpus":
        """Hands an already loaded corpus (e.g. a file-level one) to the cache; `corpus` must not be used afterwards."""
        out = C.c_void_p()
        N.check(N.lib().bsg_cache_insert(self._h, int(file_id), corpus._h, C.byref(out)))
        corpus._h = C.c_void_p()
        return self._borrow(out.value)

    def release(self, corpus: "Corpus"):
        if corpus._h:
            N.lib().bsg_cache_release(self._h, corpus._h)
            corpus._h = C.c_void_p()

    def invalidate(self, file_id: int):
        N.check(N.lib().bsg_cache_invalidate(self._h, int(file_id)))

    def stats(self) -> dict:
        v = [C.c_uint64() for _ in range(6)]
        N.check(N.lib().bsg_cache_stats(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("used_bytes", "entries", "hits", "misses", "evictions", "invalidations"), (x.value for x in v)))

    def close(self):
        if self._h:
            N.lib().bsg_cache_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def probe_hierarchical(files: Corpus, blocks: Corpus, query: Optional[BloomQuery]):
    """The reference's two-stage pruning in one call: file-level filters first (query_exec.go:399-406),
    then block-level filters only for blocks of surviving files (query_exec.go:572-615).
    Returns (file_survives bool[n_files], block_survives bool[n_blocks])."""
    cq = compile_bloom_query(query)
    blob, off = N.pack_keys(cq.keys)
    kinds = np.ascontiguousarray(cq.kinds, dtype=np.uint8)
    fmask = np.zeros((files.n_units + 63) // 64, dtype=np.uint64)
    bmask = np.zeros((blocks.n_units + 63) // 64, dtype=np.uint64)
    pp, pl = (None, 0) if cq.prog is None else (N.ptr(cq.prog), len(cq.prog))
    N.check(N.lib().bsg_probe_hierarchical(files.ctx.handle, files.handle, blocks.handle, N.ptr(blob), N.ptr(off),
                                           len(cq.keys), N.ptr(kinds), pp, pl, N.ptr(fmask), N.ptr(bmask)))
    return unpack_mask(fmask, files.n_units), unpack_mask(bmask, blocks.n_units)


class KeySet:
    """Grouped keys resident in HBM (bsg_keyset): count distinct entries, size filters on the host, build —
    the emissions cross PCIe once (ingest.go:24-145 on the device)."""

    def __init__(self, ctx: Context, blob: np.ndarray, key_off: np.ndarray, group_begin: np.ndarray):
        self.ctx = ctx
        key_off = np.ascontiguousarray(key_off, dtype=np.uint64)
        group_begin = np.ascontiguousarray(group_begin, dtype=np.uint64)
        self.n_groups = len(group_begin) - 1
        self.n_words = 0
        self._h = C.c_void_p()
        N.check(N.lib().bsg_keyset_create(ctx.handle, N.ptr(blob), N.ptr(key_off), len(key_off) - 1, N.ptr(group_begin),
                                          self.n_groups, C.byref(self._h)))

    def count_distinct(self, group_parent: Optional[np.ndarray] = None, n_parents: int = 0):
        gp = None if group_parent is None else np.ascontiguousarray(group_parent, dtype=np.uint32)
        gc = np.zeros(max(self.n_groups, 1), dtype=np.uint64)
        pc = None if gp is None else np.zeros(max(int(n_parents), 1), dtype=np.uint64)
        N.check(N.lib().bsg_keyset_count_distinct(self.ctx.handle, self._h, N.ptr(gp), int(n_parents), N.ptr(gc), N.ptr(pc)))
        return gc[:self.n_groups], (None if pc is None else pc[:int(n_parents)])

    def set_filters(self, group_filter, group_filter2, desc, n_words: int):
        desc = np.ascontiguousarray(desc, dtype=N.DESC_DTYPE)
        gf = np.ascontiguousarray(group_filter, dtype=np.uint32)
        gf2 = None if group_filter2 is None else np.ascontiguousarray(group_filter2, dtype=np.uint32)
        N.check(N.lib().bsg_keyset_set_filters(self.ctx.handle, self._h, N.ptr(gf), N.ptr(gf2), N.ptr(desc), len(desc),
                                               int(n_words)))
        self.n_words = int(n_words)

    def build(self, d_out: Optional[int] = None):
        """Asynchronous on the ctx stream; d_out = device pointer (e.g. Context.comm_alloc) or None."""
        N.check(N.lib().bsg_keyset_build(self.ctx.handle, self._h, C.c_void_p(d_out) if d_out else None))

    def fetch(self) -> np.ndarray:
        out = np.zeros(max(self.n_words, 1), dtype=np.uint64)
        N.check(N.lib().bsg_keyset_fetch(self.ctx.handle, self._h, N.ptr(out)))
        return out[:self.n_words]

    def device_words(self) -> int:
        return N.lib().bsg_keyset_device_words(self._h)

    def close(self):
        if self._h:
            N.lib().bsg_keyset_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def probe_hierarchical_gather(files: "Corpus", blocks: "Corpus", query: Optional[BloomQuery], mask_words: int, world: int):
    """Collective: every rank probes its shard hierarchically, the per-rank block masks are all-gathered on
    the device; returns uint64[world, mask_words]."""
    cq = compile_bloom_query(query)
    blob, off = N.pack_keys(cq.keys)
    kinds = np.ascontiguousarray(cq.kinds, dtype=np.uint8)
    out = np.zeros((world, mask_words), dtype=np.uint64)
    pp, pl = (None, 0) if cq.prog is None else (N.ptr(cq.prog), len(cq.prog))
    N.check(N.lib().bsg_probe_hierarchical_gather(files.ctx.handle, files.handle, blocks.handle, N.ptr(blob), N.ptr(off),
                                                  len(cq.keys), N.ptr(kinds), pp, pl, int(mask_words), N.ptr(out)))
    return out


class Query:
    """Device-resident query (keys hashed once) for repeated / timed probes."""

    def __init__(self, corpus: Corpus, keys: Sequence[bytes], kinds, prog: Optional[np.ndarray] = None):
        self.corpus = corpus
        self.n_keys = len(keys)
        blob, off = N.pack_keys(list(keys))
        kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
        pp, pl = (None, 0) if prog is None else (N.ptr(np.ascontiguousarray(prog, N.OP_DTYPE)), len(prog))
        self._h = C.c_void_p()
        N.check(N.lib().bsg_query_create(corpus.ctx.handle, corpus.handle, N.ptr(blob), N.ptr(off), self.n_keys,
                                         N.ptr(kinds), pp, pl, C.byref(self._h)))

    def run(self, path: int = N.PROBE_AUTO, corpus: Optional[Corpus] = None, want_matrix: bool = True):
        """want_matrix=False: only the candidate mask is wanted — a small query on the gather path then stops testing
        a unit's keys as soon as its expression is decided (the fetched matrix is not the membership matrix)."""
        c = corpus or self.corpus
        N.check(N.lib().bsg_query_run(c.ctx.handle, c.handle, self._h, path, 1 if want_matrix else 0))

    def run_child(self, blocks: Corpus, parent: "Query", path: int = N.PROBE_AUTO):
        """Second stage of a hierarchical probe: only units of `blocks` whose parent survived `parent`'s run."""
        N.check(N.lib().bsg_query_run_child(blocks.ctx.handle, blocks.handle, self._h, parent._h, path))

    def device_mask(self) -> int:
        return N.lib().bsg_query_device_mask(self._h)

    def launches(self) -> int:
        return N.lib().bsg_query_last_launches(self._h)

    def fetch(self, want_matrix: bool = True, want_mask: bool = True):
        n = self.corpus.n_units
        matrix = np.zeros((n, (self.n_keys + 63) // 64), dtype=np.uint64) if want_matrix else None
        mask = np.zeros((n + 63) // 64, dtype=np.uint64) if want_mask else None
        N.check(N.lib().bsg_query_fetch(self.corpus.ctx.handle, self._h, n, N.ptr(matrix), N.ptr(mask)))
        return matrix, mask

    def close(self):
        if self._h:
            N.lib().bsg_query_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def unpack_mask(mask: np.ndarray, n_units: int) -> np.ndarray:
    bits = np.unpackbits(mask.view(np.uint8), bitorder="little")
    return bits[:n_units].astype(bool)


def unpack_matrix(matrix: np.ndarray, n_keys: int) -> np.ndarray:
    """uint64[n_units, ceil(Q/64)] -> bool[n_units, Q]."""
    n_units = matrix.shape[0]
    if n_units == 0 or n_keys == 0:
        return np.zeros((n_units, n_keys), dtype=bool)
    bits = np.unpackbits(np.ascontiguousarray(matrix).view(np.uint8).reshape(n_units, -1), axis=1, bitorder="little")
    return bits[:, :n_keys].astype(bool)
