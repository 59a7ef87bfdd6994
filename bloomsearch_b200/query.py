"""Host-side mirror of the reference's bloom query AST (query.go:478-718).

Same node types, builders, flattening and rewrite helpers as the Go package, so
the parity tests read like the reference's own.  `compile_bloom_query` lowers a
tree to what the C ABI takes: a list of distinct (kind, key-bytes) leaves and a
postfix program over them (include/bloomgpu.h, bsg_expr_op).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N

# BloomConditionType (query.go:478-484)
BloomField = "FIELD"
BloomToken = "TOKEN"
BloomFieldToken = "FIELD_TOKEN"
# BloomExpressionType (query.go:493-499)
BloomExpressionCondition = "CONDITION"
BloomExpressionAnd = "AND"
BloomExpressionOr = "OR"


def _b(s) -> bytes:
    return s if isinstance(s, (bytes, bytearray)) else str(s).encode("utf-8")


@dataclass
class BloomCondition:  # query.go:486-490
    Type: str
    Field: bytes = b""
    Token: bytes = b""


@dataclass
class BloomExpression:  # query.go:505-509
    ExpressionType: str
    Condition: Optional[BloomCondition] = None
    Children: List["BloomExpression"] = field(default_factory=list)


@dataclass
class BloomQuery:  # query.go:511-513
    Expression: Optional[BloomExpression] = None


@dataclass
class RegexCondition:  # query.go:515-518
    Field: bytes
    Pattern: str = ""


@dataclass
class RegexExpression:  # query.go:533-537
    ExpressionType: str
    Condition: Optional[RegexCondition] = None
    Children: List["RegexExpression"] = field(default_factory=list)


@dataclass
class RegexQuery:
    Expression: Optional[RegexExpression] = None


def Field(field_path) -> BloomExpression:  # query.go:549-557
    return BloomExpression(BloomExpressionCondition, BloomCondition(BloomField, Field=_b(field_path)))


def Token(token) -> BloomExpression:  # query.go:559-567
    return BloomExpression(BloomExpressionCondition, BloomCondition(BloomToken, Token=_b(token)))


def FieldToken(field_path, token) -> BloomExpression:  # query.go:575-584
    return BloomExpression(BloomExpressionCondition,
                           BloomCondition(BloomFieldToken, Field=_b(field_path), Token=_b(token)))


def _flatten(expressions: Sequence[BloomExpression], expression_type: str) -> List[BloomExpression]:
    # flattenExpressions, query.go:600-610
    out: List[BloomExpression] = []
    for e in expressions:
        if e.ExpressionType == expression_type and e.Condition is None:
            out.extend(e.Children)
        else:
            out.append(e)
    return out


def And(*expressions: BloomExpression) -> BloomExpression:  # query.go:586-591
    return BloomExpression(BloomExpressionAnd, None, _flatten(expressions, BloomExpressionAnd))


def Or(*expressions: BloomExpression) -> BloomExpression:  # query.go:593-598
    return BloomExpression(BloomExpressionOr, None, _flatten(expressions, BloomExpressionOr))


def FieldRegex(field_path, pattern: str) -> RegexExpression:
    return RegexExpression("CONDITION", RegexCondition(_b(field_path), pattern))


def RegexAnd(*expressions: RegexExpression) -> RegexExpression:
    return RegexExpression("AND", None, list(expressions))


def RegexOr(*expressions: RegexExpression) -> RegexExpression:
    return RegexExpression("OR", None, list(expressions))


def regex_expression_to_bloom_field_expression(e: Optional[RegexExpression]) -> Optional[BloomExpression]:
    # regexExpressionToBloomFieldExpression, query.go:651-694
    if e is None:
        return None
    if e.ExpressionType == "CONDITION":
        if e.Condition is None:
            return None
        return BloomExpression(BloomExpressionCondition, BloomCondition(BloomField, Field=e.Condition.Field))
    if e.ExpressionType in ("AND", "OR"):
        kids = [c for c in (regex_expression_to_bloom_field_expression(ch) for ch in e.Children) if c is not None]
        return BloomExpression(BloomExpressionAnd if e.ExpressionType == "AND" else BloomExpressionOr, None, kids)
    return None


def RegexFieldGuardBloomQuery(query: Optional[RegexQuery]) -> Optional[BloomQuery]:  # query.go:696-705
    if query is None or query.Expression is None:
        return None
    e = regex_expression_to_bloom_field_expression(query.Expression)
    return None if e is None else BloomQuery(e)


def AndBloomQueries(left: Optional[BloomQuery], right: Optional[BloomQuery]) -> Optional[BloomQuery]:
    # query.go:707-716
    if left is None or left.Expression is None:
        return right
    if right is None or right.Expression is None:
        return left
    return BloomQuery(And(left.Expression, right.Expression))


def make_field_token_key(field_path: bytes, token: bytes) -> bytes:
    """makeFieldTokenKey, tokenizer.go:508-511 (must match ingest.go:95-99 byte for byte)."""
    return _b(field_path) + b"::" + _b(token)


class QueryBuilder:
    """NewQuery().Field(...).Token(...).FieldToken(...).Match(expr).Build() — the slice of
    query.go:728-833 that feeds the bloom stage (conditions AND together)."""

    def __init__(self):
        self._parts: List[BloomExpression] = []

    def Field(self, f):
        self._parts.append(Field(f))
        return self

    def Token(self, t):
        self._parts.append(Token(t))
        return self

    def FieldToken(self, f, t):
        self._parts.append(FieldToken(f, t))
        return self

    def Match(self, expr: BloomExpression):
        self._parts.append(expr)
        return self

    def Build(self) -> BloomQuery:
        if not self._parts:
            return BloomQuery(None)
        if len(self._parts) == 1:
            return BloomQuery(self._parts[0])
        return BloomQuery(And(*self._parts))


def NewQuery() -> QueryBuilder:
    return QueryBuilder()


@dataclass
class CompiledQuery:
    keys: List[bytes]          # distinct leaf keys (FieldToken already joined)
    kinds: np.ndarray          # uint8[n_keys], BSG_KIND_*
    prog: Optional[np.ndarray]  # OP_DTYPE[prog_len] or None (no expression: every unit survives)


_KIND = {BloomField: N.KIND_FIELD, BloomToken: N.KIND_TOKEN, BloomFieldToken: N.KIND_FIELDTOKEN}


def compile_bloom_query(query: Optional[BloomQuery]) -> CompiledQuery:
    """Lower a BloomQuery to (keys, kinds, postfix).  Semantics preserved from
    query_exec.go:75-159: nil query / nil expression -> no program; nil Condition ->
    TRUE; unknown expression or condition type -> FALSE; OR [] -> false; AND [] -> true."""
    if query is None or query.Expression is None:
        return CompiledQuery([], np.zeros(0, np.uint8), None)
    keys: List[bytes] = []
    kinds: List[int] = []
    index = {}
    prog: List[Tuple[int, int]] = []

    def leaf(kind: int, key: bytes) -> int:
        ix = index.get((kind, key))
        if ix is None:
            ix = len(keys)
            index[(kind, key)] = ix
            keys.append(key)
            kinds.append(kind)
        return ix

    def emit(e: Optional[BloomExpression]) -> None:
        if e is None:
            prog.append((N.OP_TRUE, 0))
            return
        if e.ExpressionType == BloomExpressionCondition:
            c = e.Condition
            if c is None:
                prog.append((N.OP_TRUE, 0))
            elif c.Type == BloomField:
                prog.append((N.OP_LEAF, leaf(N.KIND_FIELD, _b(c.Field))))
            elif c.Type == BloomToken:
                prog.append((N.OP_LEAF, leaf(N.KIND_TOKEN, _b(c.Token))))
            elif c.Type == BloomFieldToken:
                prog.append((N.OP_LEAF, leaf(N.KIND_FIELDTOKEN, make_field_token_key(c.Field, c.Token))))
            else:
                prog.append((N.OP_FALSE, 0))
        elif e.ExpressionType in (BloomExpressionAnd, BloomExpressionOr):
            op = N.OP_AND if e.ExpressionType == BloomExpressionAnd else N.OP_OR
            # n-ary nodes are folded pairwise (child, child, op 2, child, op 2, ...): a node never holds more
            # than two values, so the evaluation stack grows by one per NESTING level, not per child
            # (BSG_MAX_STACK = 64 then covers trees 63 levels deep, however wide)
            n = len(e.Children)
            if n == 0:
                prog.append((op, 0))
                return
            for i, ch in enumerate(e.Children):
                emit(ch)
                if i >= 1:
                    prog.append((op, 2))
            if n == 1:
                prog.append((op, 1))
        else:
            prog.append((N.OP_FALSE, 0))

    emit(query.Expression)
    arr = np.array(prog, dtype=np.uint32).reshape(-1, 2)
    out = np.zeros(len(prog), dtype=N.OP_DTYPE)
    out["op"] = arr[:, 0]
    out["arg"] = arr[:, 1]
    return CompiledQuery(keys, np.array(kinds, dtype=np.uint8), out)
