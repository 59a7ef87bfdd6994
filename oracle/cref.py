"""ctypes binding of oracle/_build/libbloomref.so — ORACLE, TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbloomref.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("bloomref.c", "bloomref.h")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


class Desc(C.Structure):
    _fields_ = [("m", C.c_uint64), ("k", C.c_uint64), ("word_off", C.c_uint64)]


class Op(C.Structure):
    _fields_ = [("op", C.c_uint32), ("arg", C.c_uint32)]


class Expr(C.Structure):
    pass


Expr._fields_ = [
    ("type", C.c_int32), ("has_condition", C.c_int32), ("cond_type", C.c_int32), ("n_children", C.c_int32),
    ("field", C.c_char_p), ("field_len", C.c_uint64),
    ("token", C.c_char_p), ("token_len", C.c_uint64),
    ("children", C.POINTER(Expr)),
]

DESC_DTYPE = np.dtype([("m", "<u8"), ("k", "<u8"), ("word_off", "<u8")])
OP_DTYPE = np.dtype([("op", "<u4"), ("arg", "<u4")])

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    u8p, u64p, u32p, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_void_p
    L.bref_murmur3_x64_128.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, u64p]
    L.bref_base_hashes.argtypes = [C.c_char_p, C.c_size_t, u64p]
    L.bref_location.argtypes = [u64p, C.c_uint64]
    L.bref_location.restype = C.c_uint64
    L.bref_estimate_parameters.argtypes = [C.c_uint64, C.c_double, u64p, u64p]
    L.bref_filter_new.argtypes = [C.c_uint64, C.c_uint64]
    L.bref_filter_new.restype = vp
    L.bref_filter_new_with_estimates.argtypes = [C.c_uint64, C.c_double]
    L.bref_filter_new_with_estimates.restype = vp
    L.bref_filter_free.argtypes = [vp]
    L.bref_filter_add.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.bref_filter_test.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.bref_filter_test.restype = C.c_int
    L.bref_filter_equal.argtypes = [vp, vp]
    L.bref_filter_equal.restype = C.c_int
    for name in ("m", "k", "nwords"):
        fn = getattr(L, "bref_filter_" + name)
        fn.argtypes = [vp]
        fn.restype = C.c_uint64
    L.bref_filter_words.argtypes = [vp]
    L.bref_filter_words.restype = u64p
    L.bref_filter_serialized_size.argtypes = [vp]
    L.bref_filter_serialized_size.restype = C.c_size_t
    L.bref_filter_write_to.argtypes = [vp, vp]
    L.bref_filter_write_to.restype = C.c_size_t
    L.bref_filter_read_from.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.bref_filter_read_from.restype = vp
    L.bref_build_sized_filter.argtypes = [vp, vp, C.c_uint64, C.c_double]
    L.bref_build_sized_filter.restype = vp
    L.bref_crc32c.argtypes = [C.c_char_p, C.c_size_t]
    L.bref_crc32c.restype = C.c_uint32
    L.bref_crc32c_sw.argtypes = [C.c_char_p, C.c_size_t]
    L.bref_crc32c_sw.restype = C.c_uint32
    L.bref_section_encode.argtypes = [C.POINTER(vp), vp]
    L.bref_section_encode.restype = C.c_size_t
    L.bref_section_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(vp)]
    L.bref_section_parse.restype = C.c_int
    L.bref_evaluate_bloom_filters.argtypes = [vp, vp, vp, C.POINTER(Expr)]
    L.bref_evaluate_bloom_filters.restype = C.c_int
    L.bref_eval_postfix.argtypes = [vp, C.c_uint32, vp, C.c_uint32]
    L.bref_eval_postfix.restype = C.c_int
    L.bref_build_filters.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, vp, vp, C.c_int]
    L.bref_build_filters.restype = None
    L.bref_probe_matrix.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, C.c_uint32, vp, C.c_int]
    L.bref_probe_matrix.restype = None
    L.bref_probe_mask.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_int]
    L.bref_probe_mask.restype = C.c_int
    L.bref_probe_sections.argtypes = [vp, vp, C.c_uint64, C.POINTER(Expr), vp, C.c_int]
    L.bref_probe_sections.restype = C.c_int64
    L.bref_probe_sections_matrix.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, C.c_uint32, vp, C.c_int]
    L.bref_probe_sections_matrix.restype = C.c_int64
    L.bref_encode_sections.argtypes = [vp, vp, C.c_uint64, vp, vp]
    L.bref_encode_sections.restype = C.c_uint64
    _lib = L
    return L


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def pack_keys(keys) -> tuple[np.ndarray, np.ndarray]:
    """list[bytes] -> (uint8 bytes, uint64 offsets[n+1])."""
    off = np.zeros(len(keys) + 1, dtype=np.uint64)
    if len(keys):
        off[1:] = np.cumsum([len(k) for k in keys], dtype=np.uint64)
    blob = np.frombuffer(b"".join(keys), dtype=np.uint8).copy() if len(keys) else np.zeros(0, np.uint8)
    if blob.size == 0:
        blob = np.zeros(1, np.uint8)  # keep a valid pointer
    return blob, off


def murmur3_x64_128(data: bytes, seed: int = 0):
    out = (C.c_uint64 * 2)()
    lib().bref_murmur3_x64_128(data, len(data), seed, out)
    return out[0], out[1]


def base_hashes(data: bytes):
    out = (C.c_uint64 * 4)()
    lib().bref_base_hashes(data, len(data), out)
    return tuple(out)


def location(h, i: int) -> int:
    arr = (C.c_uint64 * 4)(*h)
    return lib().bref_location(arr, i)


def estimate_parameters(n: int, p: float):
    m, k = C.c_uint64(), C.c_uint64()
    lib().bref_estimate_parameters(n, p, C.byref(m), C.byref(k))
    return m.value, k.value


def crc32c(data: bytes) -> int:
    return lib().bref_crc32c(data, len(data))


def crc32c_sw(data: bytes) -> int:
    """The table-driven path alone (bref_crc32c takes the SSE4.2 instruction where the host has it)."""
    return lib().bref_crc32c_sw(data, len(data))


class Filter:
    """Owning handle on a bref_filter."""

    def __init__(self, handle):
        if not handle:
            raise MemoryError("bref_filter allocation/decoding failed")
        self.h = handle

    @classmethod
    def new(cls, m, k):
        return cls(lib().bref_filter_new(m, k))

    @classmethod
    def with_estimates(cls, n, fpr):
        return cls(lib().bref_filter_new_with_estimates(n, fpr))

    @classmethod
    def build_sized(cls, keys, fpr):
        blob, off = pack_keys(list(keys))
        return cls(lib().bref_build_sized_filter(_p(blob), _p(off), len(off) - 1, fpr))

    @classmethod
    def read_from(cls, raw: bytes):
        used = C.c_size_t()
        return cls(lib().bref_filter_read_from(raw, len(raw), C.byref(used)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().bref_filter_free(self.h)
            self.h = None

    def add(self, data: bytes):
        lib().bref_filter_add(self.h, data, len(data))

    def test(self, data: bytes) -> bool:
        return bool(lib().bref_filter_test(self.h, data, len(data)))

    @property
    def m(self):
        return lib().bref_filter_m(self.h)

    @property
    def k(self):
        return lib().bref_filter_k(self.h)

    @property
    def nwords(self):
        return lib().bref_filter_nwords(self.h)

    def words(self) -> np.ndarray:
        n = self.nwords
        return np.ctypeslib.as_array(lib().bref_filter_words(self.h), shape=(n,)).copy()

    def write_to(self) -> bytes:
        n = lib().bref_filter_serialized_size(self.h)
        buf = C.create_string_buffer(n)
        lib().bref_filter_write_to(self.h, buf)
        return buf.raw

    def equal(self, other: "Filter") -> bool:
        return bool(lib().bref_filter_equal(self.h, other.h))


def section_encode(field, token, fieldtoken) -> bytes:
    arr = (C.c_void_p * 3)(*[f.h if f is not None else None for f in (field, token, fieldtoken)])
    n = lib().bref_section_encode(arr, None)
    buf = C.create_string_buffer(n)
    lib().bref_section_encode(arr, buf)
    return buf.raw


def section_parse(section: bytes):
    arr = (C.c_void_p * 3)()
    rc = lib().bref_section_parse(section, len(section), arr)
    if rc != 0:
        raise ValueError(f"bref_section_parse rc={rc}")
    return tuple(Filter(arr[i]) if arr[i] else None for i in range(3))


_EXPR_T = {"COND": 0, "AND": 1, "OR": 2}
_COND_T = {"FIELD": 0, "TOKEN": 1, "FIELD_TOKEN": 2}


def make_expr(expr, keep):
    """tuple-form tree (see bloomref.evaluate_bloom_filters) -> Expr; `keep` pins buffers."""
    e = Expr()
    kind = expr[0]
    e.type = _EXPR_T.get(kind, 3)
    if kind == "COND":
        cond = expr[1]
        e.has_condition = 0 if cond is None else 1
        if cond is not None:
            ctype, field, token = cond
            e.cond_type = _COND_T.get(ctype, 3)
            field = field or b""
            token = token or b""
            keep.extend([field, token])
            e.field, e.field_len = field, len(field)
            e.token, e.token_len = token, len(token)
    elif kind in ("AND", "OR"):
        kids = (Expr * max(len(expr[1]), 1))()
        for i, c in enumerate(expr[1]):
            kids[i] = make_expr(c, keep)
        keep.append(kids)
        e.n_children = len(expr[1])
        e.children = C.cast(kids, C.POINTER(Expr))
    return e


def evaluate_bloom_filters(field_f, token_f, fieldtoken_f, expr) -> bool:
    keep = []
    e = None if expr is None else C.byref(make_expr(expr, keep))
    hs = [f.h if f is not None else None for f in (field_f, token_f, fieldtoken_f)]
    return bool(lib().bref_evaluate_bloom_filters(hs[0], hs[1], hs[2], e))


def eval_postfix(prog: np.ndarray, leaf_bits: np.ndarray) -> int:
    prog = np.ascontiguousarray(prog, dtype=OP_DTYPE)
    leaf_bits = np.ascontiguousarray(leaf_bits, dtype=np.uint8)
    return lib().bref_eval_postfix(_p(prog), len(prog), _p(leaf_bits), len(leaf_bits))


def build_filters(blob, key_off, group_begin, group_filter, group_filter2, desc, n_words, n_threads=1):
    out = np.zeros(max(int(n_words), 1), dtype=np.uint64)
    gf2 = None if group_filter2 is None else _p(np.ascontiguousarray(group_filter2, np.uint32))
    group_begin = np.ascontiguousarray(group_begin, np.uint64)
    group_filter = np.ascontiguousarray(group_filter, np.uint32)
    desc = np.ascontiguousarray(desc, DESC_DTYPE)
    lib().bref_build_filters(_p(blob), _p(key_off), _p(group_begin), len(group_filter), _p(group_filter),
                             gf2, _p(desc), _p(out), n_threads)
    return out[:int(n_words)]


def probe_matrix(desc, words, n_units, blob, key_off, kinds, n_threads=1):
    q = len(key_off) - 1
    out = np.zeros((int(n_units), (q + 63) // 64), dtype=np.uint64)
    desc = np.ascontiguousarray(desc, DESC_DTYPE)
    kinds = np.ascontiguousarray(kinds, np.uint8)
    lib().bref_probe_matrix(_p(desc), _p(words), n_units, _p(blob), _p(key_off), _p(kinds), q, _p(out), n_threads)
    return out


def probe_mask(desc, words, n_units, blob, key_off, kinds, prog, n_threads=1):
    q = len(key_off) - 1
    out = np.zeros((int(n_units) + 63) // 64, dtype=np.uint64)
    desc = np.ascontiguousarray(desc, DESC_DTYPE)
    kinds = np.ascontiguousarray(kinds, np.uint8)
    if prog is None:
        pp, pl = None, 0
    else:
        prog = np.ascontiguousarray(prog, OP_DTYPE)
        pp, pl = _p(prog), len(prog)
    rc = lib().bref_probe_mask(_p(desc), _p(words), n_units, _p(blob), _p(key_off), _p(kinds), q, pp, pl,
                               _p(out), n_threads)
    if rc != 0:
        raise ValueError("malformed postfix program")
    return out


def probe_sections(sections: np.ndarray, sec_off: np.ndarray, expr, n_threads=1):
    n_units = len(sec_off) - 1
    out = np.zeros((n_units + 63) // 64, dtype=np.uint64)
    keep = []
    e = None if expr is None else C.byref(make_expr(expr, keep))
    errs = lib().bref_probe_sections(_p(sections), _p(sec_off), n_units, e, _p(out), n_threads)
    return out, errs


def probe_sections_matrix(sections: np.ndarray, sec_off: np.ndarray, blob, key_off, kinds, n_threads=1):
    n_units = len(sec_off) - 1
    q = len(key_off) - 1
    out = np.zeros((n_units, (q + 63) // 64), dtype=np.uint64)
    kinds = np.ascontiguousarray(kinds, np.uint8)
    errs = lib().bref_probe_sections_matrix(_p(sections), _p(sec_off), n_units, _p(blob), _p(key_off), _p(kinds), q,
                                            _p(out), n_threads)
    return out, errs


def encode_sections(desc, words, n_units):
    """-> (uint8 sections, uint64 sec_off[n_units+1]) in the on-disk framing."""
    desc = np.ascontiguousarray(desc, DESC_DTYPE)
    words = np.ascontiguousarray(words, np.uint64)
    sec_off = np.zeros(n_units + 1, dtype=np.uint64)
    total = lib().bref_encode_sections(_p(desc), _p(words), n_units, _p(sec_off), None)
    out = np.zeros(max(int(total), 1), dtype=np.uint8)
    lib().bref_encode_sections(_p(desc), _p(words), n_units, _p(sec_off), _p(out))
    return out[:int(total)], sec_off
