"""Synthetic benchRows-shaped corpus (synth/corpusgen.c): INPUT DATA for tests and bench.
Not part of the oracle (it computes no bloom arithmetic) and not part of the product."""
from __future__ import annotations

import ctypes as C

import numpy as np

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsynth.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "corpusgen.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else ["-s"]), stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


class SynthCorpus:
    """Entry sets of n_blocks blocks: packed keys, CSR groups (3 per block: field, token,
    fieldtoken), and each file's exact union distinct counts."""

    def __init__(self, seed: int, block_lo: int, n_blocks: int, rows_per_block: int, blocks_per_file: int):
        L = _load()
        L.bgen_generate.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.bgen_generate.restype = C.c_void_p
        L.bgen_free.argtypes = [C.c_void_p]
        for name, rt in (("bgen_bytes", C.POINTER(C.c_uint8)), ("bgen_key_off", C.POINTER(C.c_uint64)),
                         ("bgen_group_begin", C.POINTER(C.c_uint64)), ("bgen_file_counts", C.POINTER(C.c_uint64)),
                         ("bgen_n_keys", C.c_uint64), ("bgen_n_files", C.c_uint64)):
            fn = getattr(L, name)
            fn.argtypes = [C.c_void_p]
            fn.restype = rt
        h = L.bgen_generate(seed, block_lo, n_blocks, rows_per_block, blocks_per_file)
        if not h:
            raise ValueError("bgen_generate failed (n_blocks must be a multiple of blocks_per_file)")
        try:
            self.n_blocks = n_blocks
            self.rows_per_block = rows_per_block
            self.blocks_per_file = blocks_per_file
            self.n_keys = L.bgen_n_keys(h)
            self.n_files = L.bgen_n_files(h)
            self.key_off = np.ctypeslib.as_array(L.bgen_key_off(h), shape=(self.n_keys + 1,)).copy()
            nbytes = int(self.key_off[-1])
            self.blob = (np.ctypeslib.as_array(L.bgen_bytes(h), shape=(max(nbytes, 1),)).copy()
                         if nbytes else np.zeros(1, np.uint8))
            self.group_begin = np.ctypeslib.as_array(L.bgen_group_begin(h), shape=(3 * n_blocks + 1,)).copy()
            self.file_counts = np.ctypeslib.as_array(L.bgen_file_counts(h), shape=(max(self.n_files, 1), 3)).copy()
        finally:
            L.bgen_free(h)

    def group_counts(self) -> np.ndarray:
        return np.diff(self.group_begin).reshape(self.n_blocks, 3)

    def key(self, i: int) -> bytes:
        return self.blob[int(self.key_off[i]):int(self.key_off[i + 1])].tobytes()

    def group_keys(self, block: int, kind: int) -> list:
        g = 3 * block + kind
        return [self.key(i) for i in range(int(self.group_begin[g]), int(self.group_begin[g + 1]))]
