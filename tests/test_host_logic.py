"""CPU tests of the host-side mirror: query AST builders / rewrites (query.go:549-718) and
the lowering of a BloomExpression to the ABI's postfix program."""
from __future__ import annotations

import itertools
import random

import numpy as np

import bloomsearch_b200 as bs
from bloomsearch_b200 import _native as N
from bloomsearch_b200.query import BloomCondition, BloomExpression, RegexExpression, RegexQuery
from tests.helpers import to_oracle_tuple
from oracle import bloomref as py
from oracle import cref


def test_and_or_flatten_like_the_reference():
    # flattenExpressions, query.go:600-610
    e = bs.And(bs.Field("a"), bs.And(bs.Field("b"), bs.Token("c")), bs.Or(bs.Field("d"), bs.Field("e")))
    assert e.ExpressionType == "AND" and len(e.Children) == 4
    assert [c.ExpressionType for c in e.Children] == ["CONDITION", "CONDITION", "CONDITION", "OR"]
    o = bs.Or(bs.Or(bs.Field("a"), bs.Field("b")), bs.Field("c"))
    assert len(o.Children) == 3


def test_field_token_key_layout():
    assert bs.make_field_token_key(b"service", b"auth") == b"service::auth"  # tokenizer.go:508-511
    cq = bs.compile_bloom_query(bs.NewQuery().FieldToken("nested.region", "region-3").Build())
    assert cq.keys == [b"nested.region::region-3"] and list(cq.kinds) == [N.KIND_FIELDTOKEN]


def test_query_builder_ands_conditions():
    q = bs.NewQuery().Field("a").Token("b").Build()
    assert q.Expression.ExpressionType == "AND" and len(q.Expression.Children) == 2
    assert bs.NewQuery().Build().Expression is None
    assert bs.compile_bloom_query(bs.NewQuery().Build()).prog is None
    assert bs.compile_bloom_query(None).prog is None


def test_and_bloom_queries_and_regex_guard():
    # query.go:696-716
    left = bs.BloomQuery(bs.Token("x"))
    assert bs.AndBloomQueries(None, left) is left
    assert bs.AndBloomQueries(left, bs.BloomQuery(None)) is left
    guard = bs.RegexFieldGuardBloomQuery(RegexQuery(bs.RegexAnd(bs.FieldRegex("msg", "a.*"),
                                                                 bs.RegexOr(bs.FieldRegex("lvl", "e"), bs.FieldRegex("svc", "p")))))
    assert guard.Expression.ExpressionType == "AND"
    assert guard.Expression.Children[0].Condition.Type == "FIELD"
    assert guard.Expression.Children[1].ExpressionType == "OR"
    both = bs.AndBloomQueries(left, guard)
    assert both.Expression.ExpressionType == "AND" and len(both.Expression.Children) == 3  # flattened
    assert bs.RegexFieldGuardBloomQuery(None) is None
    assert bs.RegexFieldGuardBloomQuery(RegexQuery(RegexExpression("BOGUS"))) is None


def test_leaf_dedup_and_kinds():
    e = bs.And(bs.Token("a"), bs.Or(bs.Token("a"), bs.Field("a")), bs.FieldToken("f", "a"))
    cq = bs.compile_bloom_query(bs.BloomQuery(e))
    assert cq.keys == [b"a", b"a", b"f::a"]
    assert list(cq.kinds) == [1, 0, 2]


def _rand_tree(rng, depth, leaves):
    r = rng.random()
    if depth == 0 or r < 0.35:
        c = rng.random()
        if c < 0.05:
            return BloomExpression("CONDITION", None)
        if c < 0.08:
            return BloomExpression("CONDITION", BloomCondition("BOGUS", b"x", b"y"))
        if c < 0.1:
            return BloomExpression("BOGUS")
        kind, key = rng.choice(leaves)
        if kind == 0:
            return bs.Field(key)
        if kind == 1:
            return bs.Token(key)
        return bs.FieldToken(key, b"t")
    n = rng.choice([0, 1, 2, 2, 3, 5])
    kids = [_rand_tree(rng, depth - 1, leaves) for _ in range(n)]
    return BloomExpression(rng.choice(["AND", "OR"]), None, kids)  # unflattened on purpose


def test_postfix_lowering_equals_recursive_evaluation():
    """compile_bloom_query + bref_eval_postfix == the recursive evaluateBloomExpression for
    random trees over random leaf truth assignments (filters replaced by a truth table)."""
    rng = random.Random(42)
    leaves = [(k, b"k%d" % i) for i, k in enumerate([0, 1, 2, 0, 1, 2, 1])]

    class Fake:
        def __init__(self, truth):
            self.truth = truth

        def test(self, key):
            return self.truth[key]

    for _ in range(300):
        tree = _rand_tree(rng, 4, leaves)
        cq = bs.compile_bloom_query(bs.BloomQuery(tree))
        assert cq.prog is not None
        for _ in range(6):
            truth = {}
            bits = []
            for kind, key in zip(cq.kinds, cq.keys):
                v = rng.random() < 0.5
                truth[(int(kind), key)] = v
                bits.append(v)
            fakes = [Fake({k: v for (kd, k), v in truth.items() if kd == i}) for i in range(3)]
            # keys never referenced by this tree default to False
            for f in fakes:
                f.truth = __import__("collections").defaultdict(bool, f.truth)
            want = py.evaluate_bloom_filters(fakes[0], fakes[1], fakes[2], to_oracle_tuple(tree))
            got = cref.eval_postfix(cq.prog, np.array(bits, dtype=np.uint8))
            assert got == int(want)


def test_wide_nodes_are_folded_to_bound_stack_depth():
    kids = [bs.Token("t%d" % i) for i in range(200)]
    cq = bs.compile_bloom_query(bs.BloomQuery(bs.Or(*kids)))
    sp = mx = 0
    for op, arg in cq.prog:
        sp = sp + 1 if op in (N.OP_LEAF, N.OP_TRUE, N.OP_FALSE) else sp - arg + 1
        mx = max(mx, sp)
    assert sp == 1 and mx <= N.MAX_STACK

    for hit in (None, 0, 57, 199):
        bits = np.zeros(200, np.uint8)
        if hit is not None:
            bits[hit] = 1
        assert cref.eval_postfix(cq.prog, bits) == int(hit is not None)
    cq = bs.compile_bloom_query(bs.BloomQuery(bs.And(*kids)))
    assert cref.eval_postfix(cq.prog, np.ones(200, np.uint8)) == 1
    bits = np.ones(200, np.uint8)
    bits[131] = 0
    assert cref.eval_postfix(cq.prog, bits) == 0


def test_entry_sets_mirror():
    a, b = bs.BloomEntrySets(), bs.BloomEntrySets()
    a.add_field(b"f")
    a.add_token(b"t")
    a.add_field_token(b"f", b"t")
    b.add_token(b"t")
    b.add_token(b"u")
    dst = bs.BloomEntrySets()
    a.union_into(dst)
    b.union_into(dst)
    assert dst.counts() == {"Fields": 1, "Tokens": 2, "FieldTokens": 1}
    assert b"f::t" in dst.fieldTokens


def test_mask_and_matrix_unpack():
    m = np.array([[0b1011, 0], [1 << 63, 1]], dtype=np.uint64)
    bits = bs.unpack_matrix(m, 70)
    assert bits.shape == (2, 70)
    assert list(np.nonzero(bits[0])[0]) == [0, 1, 3] and list(np.nonzero(bits[1])[0]) == [63, 64]
    assert list(bs.unpack_mask(np.array([0b101], dtype=np.uint64), 3)) == [True, False, True]
    # nested wide nodes: the stack grows by one per nesting level, not per child (pairwise fold)
    inner = bs.And(*[bs.Token("a%d" % i) for i in range(40)])
    mid = bs.Or(*([bs.Field("f%d" % i) for i in range(39)] + [inner]))
    outer = bs.And(*([bs.Token("b%d" % i) for i in range(50)] + [mid]))
    cq2 = bs.compile_bloom_query(bs.BloomQuery(outer))
    sp = mx = 0
    for op, arg in cq2.prog:
        sp = sp + 1 if op in (N.OP_LEAF, N.OP_TRUE, N.OP_FALSE) else sp - arg + 1
        mx = max(mx, sp)
    assert sp == 1 and mx <= 4
    bits = np.ones(len(cq2.keys), np.uint8)
    assert cref.eval_postfix(cq2.prog, bits) == 1
    bits[cq2.keys.index(b"a7")] = 0            # one inner AND leaf false, but the OR's fields are true
    assert cref.eval_postfix(cq2.prog, bits) == 1
    bits[:] = 1
    bits[cq2.keys.index(b"b3")] = 0
    assert cref.eval_postfix(cq2.prog, bits) == 0


def test_mod_m32_formula(tmp_path):
    """The exact 32-bit `location % m` of the kernels (bsg_device.cuh: mod_m32), same operations in
    plain C, against 128-bit arithmetic on edge moduli and edge values (tools/mod32_check.c)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "mod32_check"
    subprocess.run(["gcc", "-O2", "-o", str(exe), os.path.join(root, "tools", "mod32_check.c")], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.strip().endswith(", 0 bad"), out
    # the C file must state the formula the device header uses (guards against the two drifting apart)
    dev = open(os.path.join(root, "bloomsearch_b200", "csrc", "bsg_device.cuh")).read()
    assert "mad.wide.u32 s, %3, %4, s" in dev and "xh * ih + shi" in dev
