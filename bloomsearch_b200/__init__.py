"""bloomsearch_b200 — B200 (sm_100a) implementation of bloomsearch's bloom-filter
build / probe hot path behind a C ABI (include/bloomgpu.h).

Python here is a thin host: ctypes over libbloomgpu.so plus a mirror of the
reference's query AST and call sites so parity tests read like the reference's.
There is no CPU fallback: without the built .so or without a CUDA device every
operation raises.
"""
from . import _native
from ._native import BloomGpuError, build
from .engine import (Batcher, BloomEntrySets, BloomFilter, BloomFilters, Context, Corpus, FilterCache, KeySet, Query, build_filters_many,
                     estimate_parameters, probe_hierarchical, probe_hierarchical_gather, unpack_mask, unpack_matrix)
from .query import (And, AndBloomQueries, BloomCondition, BloomExpression, BloomQuery, Field, FieldRegex, FieldToken,
                    NewQuery, Or, RegexAnd, RegexFieldGuardBloomQuery, RegexOr, RegexQuery, Token,
                    compile_bloom_query, make_field_token_key)

__all__ = [n for n in dir() if not n.startswith("_")]
