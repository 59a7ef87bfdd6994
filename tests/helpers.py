"""Shared test helpers: seeded inputs and oracle-side reference computations."""
from __future__ import annotations

import random

import numpy as np

from oracle import cref


def rand_keys(rng: random.Random, n: int, min_len=0, max_len=40, alphabet=None):
    out = set()
    while len(out) < n:
        L = rng.randint(min_len, max_len)
        if alphabet:
            out.add(bytes(rng.choice(alphabet) for _ in range(L)))
        else:
            out.add(bytes(rng.randrange(256) for _ in range(L)))
    return sorted(out)


def oracle_units(unit_keys, fpr, absent=()):
    """unit_keys: list of (fields, tokens, fieldtokens) key lists -> (desc, words) built by the
    C oracle exactly like buildFilters (ingest.go:127-145).  absent: set of (unit, kind) slots
    to leave as nil filters."""
    desc = np.zeros(len(unit_keys) * 3, dtype=cref.DESC_DTYPE)
    chunks, off = [], 0
    for u, kinds in enumerate(unit_keys):
        for kind, keys in enumerate(kinds):
            if (u, kind) in absent:
                continue
            f = cref.Filter.build_sized(keys, fpr)
            w = f.words()
            desc[u * 3 + kind] = (f.m, f.k, off)
            chunks.append(w)
            off += len(w)
    words = np.concatenate(chunks) if chunks else np.zeros(0, np.uint64)
    return desc, words


def to_oracle_tuple(e):
    """bloomsearch_b200.query.BloomExpression -> the tuple form the oracle evaluators take."""
    if e is None:
        return None
    b = lambda x: x if isinstance(x, (bytes, bytearray)) else str(x).encode()
    if e.ExpressionType == "CONDITION":
        if e.Condition is None:
            return ("COND", None)
        c = e.Condition
        return ("COND", (c.Type, b(c.Field), b(c.Token)))
    if e.ExpressionType in ("AND", "OR"):
        return (e.ExpressionType, [to_oracle_tuple(c) for c in e.Children])
    return (e.ExpressionType, [])
