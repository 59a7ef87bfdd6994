"""ctypes bindings of the third-party pins under oracle/_ref/ — TEST INFRASTRUCTURE ONLY.

(1) libmurmur3_canonical.so, below; (2) libcrc32c_hw.so: CRC32C computed by the CPU's SSE4.2 `crc32`
instruction (oracle/pins/crc32c_hw.c), against which the oracle's table-driven CRC32C (bref_crc32c_sw) and the Python twin are checked — the
checksum of every filter section (/root/reference/file_format.go:44,379,399).

The library is Austin Appleby's canonical MurmurHash3.cpp (public domain; the SMHasher source),
compiled UNMODIFIED from where scikit-learn ships it in this image (sklearn/utils/src/, see
oracle/Makefile target `ref`); no source is copied into this repository.  It is not the reference
(bloomsearch is Go) and not bloom/v3; it is the algorithm bloom/v3 v3.7.0's murmur.go states its
`sum256` is strictly equivalent to:

    hasher := murmur3.New128(); hasher.Write(data); v1, v2 := hasher.Sum128()
    hasher.Write([]byte{1});                        v3, v4 := hasher.Sum128()

(/root/reference/ingest.go:139-145 and query_exec.go:128-159 reach it through AddString /
TestString -> baseHashes).  tests/ use it to check oracle/bloomref.c, oracle/bloomref.py and the
CUDA hash (csrc/bsg_device.cuh) for every key length and tail shape.  Only tests/ may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libmurmur3_canonical.so")
_lib = None


def build() -> str | None:
    """Compiles the pin if scikit-learn's copy of MurmurHash3.cpp is present; returns the .so path or None."""
    if not os.path.exists(_SO):
        subprocess.call(["make", "-s", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO if os.path.exists(_SO) else None


def available() -> bool:
    return build() is not None


def lib():
    global _lib
    if _lib is None:
        so = build()
        if so is None:
            raise RuntimeError("canonical MurmurHash3.cpp not found (scikit-learn's sklearn/utils/src/ is absent)")
        _lib = C.CDLL(so)
        _lib.MurmurHash3_x64_128.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.POINTER(C.c_uint64)]
        _lib.MurmurHash3_x64_128.restype = None
    return _lib


def murmur3_x64_128(data: bytes, seed: int = 0) -> tuple[int, int]:
    out = (C.c_uint64 * 2)()
    lib().MurmurHash3_x64_128(data, len(data), seed, out)
    return int(out[0]), int(out[1])


def base_hashes(data: bytes) -> tuple[int, int, int, int]:
    """What bloom/v3 documents baseHashes(data) to equal, computed ONLY with the canonical code."""
    return murmur3_x64_128(data) + murmur3_x64_128(data + b"\x01")


# ---- second pin: CRC32C by the CPU's own SSE4.2 instruction (oracle/pins/crc32c_hw.c) ----
_SO_CRC = os.path.join(_HERE, "_ref", "libcrc32c_hw.so")
_lib_crc = None


def build_crc32c_hw() -> str | None:
    """Compiles the hardware CRC32C pin on x86-64 hosts with SSE4.2; returns the .so path or None."""
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        flags = ""
    if "sse4_2" not in flags:
        return None
    if not os.path.exists(_SO_CRC):
        subprocess.call(["make", "-s", "-C", _HERE, "ref-crc"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO_CRC if os.path.exists(_SO_CRC) else None


def crc32c_hw_available() -> bool:
    return build_crc32c_hw() is not None


def crc32c_hw(data: bytes) -> int:
    global _lib_crc
    if _lib_crc is None:
        so = build_crc32c_hw()
        if so is None:
            raise RuntimeError("no SSE4.2 crc32 instruction on this host")
        _lib_crc = C.CDLL(so)
        _lib_crc.crc32c_hw.argtypes = [C.c_char_p, C.c_size_t]
        _lib_crc.crc32c_hw.restype = C.c_uint32
    return int(_lib_crc.crc32c_hw(data, len(data)))
