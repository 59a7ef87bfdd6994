"""ctypes binding of libbloomgpu.so (include/bloomgpu.h).

The library is built in-tree (bloomsearch_b200/_build/libbloomgpu.so) by
`build()`; there is no fallback of any kind — if the shared object is missing or
no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(_HERE, "_build", "libbloomgpu.so")

BSG_OK = 0
ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_FORMAT, ERR_UNSUPPORTED, ERR_COMM = -1, -2, -3, -4, -5, -6
KIND_FIELD, KIND_TOKEN, KIND_FIELDTOKEN = 0, 1, 2
OP_LEAF, OP_AND, OP_OR, OP_TRUE, OP_FALSE = 0, 1, 2, 3, 4
PROBE_AUTO, PROBE_STAGED, PROBE_GATHER = 0, 1, 2
RUN_MATRIX_ONLY = 0x100
NO_FILTER = 0xFFFFFFFF
MAX_STACK = 64

DESC_DTYPE = np.dtype([("m", "<u8"), ("k", "<u8"), ("word_off", "<u8")])
OP_DTYPE = np.dtype([("op", "<u4"), ("arg", "<u4")])


class BloomGpuError(RuntimeError):
    def __init__(self, code: int, detail: str):
        super().__init__(f"libbloomgpu error {code}: {detail}")
        self.code = code
        self.detail = detail


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libbloomgpu.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    host = os.path.join(_HERE, "host")
    srcs = ([os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(host, f) for f in os.listdir(host)] +
            [os.path.join(_HERE, "..", "include", "bloomgpu.h")])
    newest_out = min((os.path.getmtime(p) if os.path.exists(p) else 0.0) for p in
                     (SO_PATH, os.path.join(_HERE, "_build", "libbloomsearch_host.so"),
                      os.path.join(_HERE, "_build", "host_selftest")))
    stale = newest_out == 0.0 or any(os.path.getmtime(s) > newest_out for s in srcs)
    if force or stale:
        cmd = ["make", "-C", CSRC] + (["-B"] if force else [])
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or verbose:
            print(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("building libbloomgpu.so failed")
    return SO_PATH


_lib = None

# every symbol include/bloomgpu.h declares (tests check the .so exports them all)
ABI_SYMBOLS = [
    "bsg_abi_version", "bsg_strerror", "bsg_last_error", "bsg_create", "bsg_destroy", "bsg_set_stream",
    "bsg_synchronize", "bsg_device_info", "bsg_estimate", "bsg_hash_keys", "bsg_build", "bsg_build_fieldtokens", "bsg_count_distinct", "bsg_corpus_load",
    "bsg_corpus_load_sections", "bsg_corpus_free", "bsg_corpus_units", "bsg_corpus_bitset_bytes",
    "bsg_corpus_unit_desc", "bsg_corpus_set_parents", "bsg_probe_hierarchical", "bsg_probe", "bsg_query_create", "bsg_query_run", "bsg_query_fetch", "bsg_query_free",
    "bsg_query_last_launches", "bsg_timer_begin", "bsg_timer_end", "bsg_comm_unique_id", "bsg_comm_init",
    "bsg_or_reduce", "bsg_allgather_masks",
    "bsg_keyset_create", "bsg_keyset_count_distinct", "bsg_keyset_set_filters", "bsg_keyset_build", "bsg_keyset_fetch",
    "bsg_keyset_device_words", "bsg_keyset_free", "bsg_query_run_child", "bsg_query_device_mask", "bsg_comm_info",
    "bsg_comm_alloc", "bsg_comm_free", "bsg_or_reduce_device", "bsg_allgather_masks_device", "bsg_probe_hierarchical_gather",
    "bsg_corpus_device_bytes", "bsg_cache_create", "bsg_cache_destroy", "bsg_cache_acquire", "bsg_cache_insert",
    "bsg_cache_insert_sections", "bsg_cache_release", "bsg_cache_invalidate", "bsg_cache_stats",
    "bsg_host_alloc", "bsg_host_free",
    "bsg_probe_multi", "bsg_batcher_create", "bsg_batcher_destroy", "bsg_batcher_probe", "bsg_batcher_stats",
]


def lib():
    """Load the shared library (never builds implicitly on import of a stale tree: call build() first)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise FileNotFoundError(
            f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(bloomsearch_b200 has no CPU fallback)")
    L = C.CDLL(SO_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    L.bsg_abi_version.restype = i32
    L.bsg_strerror.argtypes = [i32]
    L.bsg_strerror.restype = C.c_char_p
    L.bsg_last_error.restype = C.c_char_p
    L.bsg_create.argtypes = [i32, C.POINTER(vp)]
    L.bsg_destroy.argtypes = [vp]
    L.bsg_destroy.restype = None
    L.bsg_set_stream.argtypes = [vp, vp]
    L.bsg_synchronize.argtypes = [vp]
    L.bsg_device_info.argtypes = [vp, C.POINTER(i32), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                  C.POINTER(i32), C.POINTER(i32)]
    L.bsg_estimate.argtypes = [u64, C.c_double, C.POINTER(u64), C.POINTER(u64)]
    L.bsg_estimate.restype = None
    L.bsg_hash_keys.argtypes = [vp, vp, vp, u64, vp]
    L.bsg_build.argtypes = [vp, vp, vp, u64, vp, u32, vp, vp, vp, u32, vp, u64]
    L.bsg_build_fieldtokens.argtypes = [vp, vp, vp, u64, vp, vp, u64, vp, u32, vp, vp, vp, u32, vp, u64]
    L.bsg_count_distinct.argtypes = [vp, vp, vp, u64, vp, u32, vp, u32, vp, vp]
    L.bsg_corpus_load.argtypes = [vp, vp, u64, vp, u64, i32, C.POINTER(vp)]
    L.bsg_corpus_load_sections.argtypes = [vp, vp, vp, u64, i32, vp, C.POINTER(u64), C.POINTER(vp)]
    L.bsg_corpus_free.argtypes = [vp]
    L.bsg_corpus_free.restype = None
    L.bsg_corpus_units.argtypes = [vp]
    L.bsg_corpus_units.restype = u64
    L.bsg_corpus_bitset_bytes.argtypes = [vp, u32]
    L.bsg_corpus_bitset_bytes.restype = u64
    L.bsg_corpus_unit_desc.argtypes = [vp, u64, vp]
    L.bsg_probe.argtypes = [vp, vp, vp, vp, u32, vp, vp, u32, vp, vp]
    L.bsg_corpus_set_parents.argtypes = [vp, vp, vp, u64, u64]
    L.bsg_probe_hierarchical.argtypes = [vp, vp, vp, vp, vp, u32, vp, vp, u32, vp, vp]
    L.bsg_query_create.argtypes = [vp, vp, vp, vp, u32, vp, vp, u32, C.POINTER(vp)]
    L.bsg_query_run.argtypes = [vp, vp, vp, i32, i32]
    L.bsg_query_fetch.argtypes = [vp, vp, u64, vp, vp]
    L.bsg_query_free.argtypes = [vp]
    L.bsg_query_free.restype = None
    L.bsg_query_last_launches.argtypes = [vp]
    L.bsg_timer_begin.argtypes = [vp]
    L.bsg_timer_end.argtypes = [vp, C.POINTER(C.c_float)]
    L.bsg_comm_unique_id.argtypes = [vp]
    L.bsg_comm_init.argtypes = [vp, i32, i32, vp]
    L.bsg_or_reduce.argtypes = [vp, vp, u64]
    L.bsg_allgather_masks.argtypes = [vp, vp, u64, vp]
    L.bsg_keyset_create.argtypes = [vp, vp, vp, u64, vp, u32, C.POINTER(vp)]
    L.bsg_keyset_count_distinct.argtypes = [vp, vp, vp, u32, vp, vp]
    L.bsg_keyset_set_filters.argtypes = [vp, vp, vp, vp, vp, u32, u64]
    L.bsg_keyset_build.argtypes = [vp, vp, vp]
    L.bsg_keyset_fetch.argtypes = [vp, vp, vp]
    L.bsg_keyset_device_words.argtypes = [vp]
    L.bsg_keyset_device_words.restype = vp
    L.bsg_keyset_free.argtypes = [vp]
    L.bsg_keyset_free.restype = None
    L.bsg_query_run_child.argtypes = [vp, vp, vp, vp, i32]
    L.bsg_query_device_mask.argtypes = [vp]
    L.bsg_query_device_mask.restype = vp
    L.bsg_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(u64)]
    L.bsg_comm_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.bsg_comm_free.argtypes = [vp, vp]
    L.bsg_or_reduce_device.argtypes = [vp, vp, u64]
    L.bsg_allgather_masks_device.argtypes = [vp, vp, u64, vp]
    L.bsg_probe_hierarchical_gather.argtypes = [vp, vp, vp, vp, vp, u32, vp, vp, u32, u64, vp]
    L.bsg_corpus_device_bytes.argtypes = [vp]
    L.bsg_corpus_device_bytes.restype = u64
    L.bsg_cache_create.argtypes = [vp, u64, C.POINTER(vp)]
    L.bsg_cache_destroy.argtypes = [vp]
    L.bsg_cache_destroy.restype = None
    L.bsg_cache_acquire.argtypes = [vp, u64, C.POINTER(vp)]
    L.bsg_cache_insert.argtypes = [vp, u64, vp, C.POINTER(vp)]
    L.bsg_cache_insert_sections.argtypes = [vp, u64, vp, vp, u64, i32, vp, C.POINTER(u64), C.POINTER(vp)]
    L.bsg_cache_release.argtypes = [vp, vp]
    L.bsg_cache_release.restype = None
    L.bsg_cache_invalidate.argtypes = [vp, u64]
    L.bsg_cache_stats.argtypes = [vp] + [C.POINTER(u64)] * 6
    L.bsg_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.bsg_host_free.argtypes = [vp, vp]
    L.bsg_probe_multi.argtypes = [vp, vp, vp, vp, u32, vp, u32, vp, vp, vp, vp]
    L.bsg_batcher_create.argtypes = [vp, vp, u32, u32, u32, C.POINTER(vp)]
    L.bsg_batcher_destroy.argtypes = [vp]
    L.bsg_batcher_destroy.restype = None
    L.bsg_batcher_probe.argtypes = [vp, vp, vp, u32, vp, vp, u32, vp]
    L.bsg_batcher_stats.argtypes = [vp] + [C.POINTER(u64)] * 4
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != BSG_OK:
        raise BloomGpuError(rc, lib().bsg_last_error().decode(errors="replace"))


def ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def pack_keys(keys) -> tuple[np.ndarray, np.ndarray]:
    """list[bytes] -> (uint8 blob, uint64 offsets[n+1]) in the ABI's packed-key layout."""
    n = len(keys)
    off = np.zeros(n + 1, dtype=np.uint64)
    if n:
        off[1:] = np.cumsum(np.fromiter((len(k) for k in keys), dtype=np.uint64, count=n))
    blob = np.frombuffer(b"".join(keys), dtype=np.uint8).copy() if n else np.zeros(0, np.uint8)
    if blob.size == 0:
        blob = np.zeros(1, np.uint8)
    return blob, off
