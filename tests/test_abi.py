"""CPU tests of the C-ABI boundary: the library loads, exports exactly what
include/bloomgpu.h declares, and fails loudly (never silently falls back) without a GPU."""
from __future__ import annotations

import ctypes as C
import os
import re

import pytest

import bloomsearch_b200 as bs
from bloomsearch_b200 import _native as N
from oracle import cref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "bloomgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(bsg_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = N.lib()
    declared = _declared_symbols()
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(N.ABI_SYMBOLS) == declared  # the Python binding tracks the header


def test_abi_version_and_strerror():
    lib = N.lib()
    assert lib.bsg_abi_version() == 2
    assert lib.bsg_strerror(0) == b"ok"
    assert b"invalid" in lib.bsg_strerror(-1)
    assert b"CUDA" in lib.bsg_strerror(-2)


def test_header_cites_reference_call_sites():
    hdr = open(os.path.join(ROOT, "include", "bloomgpu.h")).read()
    for cite in ("ingest.go:127-145", "query_exec.go:75-159", "file_format.go:392-448", "flush.go:204,253"):
        assert cite in hdr


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bs.BloomGpuError) as ei:
        bs.Context(0)
    assert ei.value.code == N.ERR_CUDA
    assert "no CPU fallback" in ei.value.detail


def test_null_arguments_are_rejected_not_crashing():
    lib = N.lib()
    assert lib.bsg_create(0, None) == N.ERR_INVALID
    assert lib.bsg_synchronize(None) == N.ERR_INVALID
    assert lib.bsg_probe(None, None, None, None, 0, None, None, 0, None, None) == N.ERR_INVALID
    assert lib.bsg_build(None, None, None, 0, None, 0, None, None, None, 0, None, 0) == N.ERR_INVALID
    assert lib.bsg_corpus_units(None) == 0
    lib.bsg_corpus_free(None)
    lib.bsg_query_free(None)
    lib.bsg_destroy(None)


def test_bsg_estimate_matches_oracle():
    for n in (1, 2, 9, 100, 101, 220, 1000, 19557, 10 ** 6, 10 ** 8):
        for p in (0.5, 0.02, 0.01, 0.001, 1e-9):
            assert bs.estimate_parameters(n, p) == cref.estimate_parameters(n, p)


def test_product_does_not_import_the_oracle():
    """The product package must never route through oracle/ (tier rule)."""
    pkg = os.path.join(ROOT, "bloomsearch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "bloomref" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_go_shim_and_cpp_host_bind_only_declared_entry_points():
    """The cgo shim (go/bloomgpu, source only: no Go toolchain here) and the C++ host layer must call
    nothing but what include/bloomgpu.h declares, and the shim must cover the whole hot path
    (build, fused field::token build, distinct counts, corpus loads, probe, hierarchical probe)."""
    declared = set(_declared_symbols())
    types = {"bsg_ctx", "bsg_corpus", "bsg_query", "bsg_expr_op", "bsg_filter_desc", "bsg_cache", "bsg_keyset", "bsg_batcher"}
    go_calls = set()
    for fn in os.listdir(os.path.join(ROOT, "go", "bloomgpu")):
        if fn.endswith(".go"):
            go_calls |= set(re.findall(r"\bC\.(bsg_[a-z_0-9]+)", open(os.path.join(ROOT, "go", "bloomgpu", fn)).read()))
    go_funcs = go_calls - types
    assert go_funcs and go_funcs <= declared, sorted(go_funcs - declared)
    for need in ("bsg_create", "bsg_destroy", "bsg_build", "bsg_build_fieldtokens", "bsg_count_distinct",
                 "bsg_corpus_load", "bsg_corpus_load_sections", "bsg_probe", "bsg_probe_hierarchical", "bsg_last_error"):
        assert need in go_funcs, need
    host_dir = os.path.join(ROOT, "bloomsearch_b200", "host")
    cpp_calls = set()
    for fn in os.listdir(host_dir):
        if fn.endswith((".cpp", ".hpp")):
            cpp_calls |= set(re.findall(r"\b(bsg_[a-z_0-9]+)\s*\(", open(os.path.join(host_dir, fn)).read()))
    assert cpp_calls and cpp_calls <= declared, sorted(cpp_calls - declared)


def _declared_arity():
    """name -> number of parameters, from include/bloomgpu.h."""
    hdr = open(os.path.join(ROOT, "include", "bloomgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(bsg_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def _call_arities(src, prefix):
    """(name, number of arguments) of every `<prefix>bsg_xxx(...)` call in src (top-level commas)."""
    out = []
    for m in re.finditer(prefix + r"(bsg_[a-z_0-9]+)\s*\(", src):
        i, depth, commas, empty = m.end(), 1, 0, True
        while depth and i < len(src):
            c = src[i]
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
            elif c == "," and depth == 1:
                commas += 1
            if depth and not c.isspace():
                empty = False
            i += 1
        out.append((m.group(1), 0 if empty else commas + 1))
    return out


def test_bindings_pass_as_many_arguments_as_the_header_declares():
    """Signature drift check for the bindings nothing here compiles or type-checks: every cgo call in
    go/bloomgpu (no Go toolchain in the image) and every ctypes argtypes list in _native.py must have exactly
    the number of parameters include/bloomgpu.h declares for that entry point."""
    decl = _declared_arity()
    assert len(decl) >= 60
    bad = []
    n_calls = 0
    for fn in sorted(os.listdir(os.path.join(ROOT, "go", "bloomgpu"))):
        if fn.endswith(".go"):
            src = re.sub(r"//.*", "", open(os.path.join(ROOT, "go", "bloomgpu", fn)).read())
            for name, n in _call_arities(src, r"\bC\."):
                if name in decl:
                    n_calls += 1
                    if decl[name] != n:
                        bad.append((fn, name, n, decl[name]))
    assert n_calls >= 20 and not bad, bad
    lib = N.lib()
    for name, n in decl.items():
        fn = getattr(lib, name)
        if fn.argtypes is not None:
            assert len(fn.argtypes) == n, (name, len(fn.argtypes), n)
