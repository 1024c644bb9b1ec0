"""Tiny SQL front end for the query shapes the reference's own tests use on this path.

The reference compiles SQL with Calcite in pinot-common (``CalciteSqlParser`` -> ``PinotQuery`` ->
``QueryContextConverterUtils.getQueryContext``), which is OUT OF SCOPE here (the broker/server path above the plan maker
stays Java).  This module exists so the parity tests read like the reference's
(``getOperator("SELECT COUNT(*), SUM(column1) FROM testTable WHERE ... GROUP BY column9")``,
``pinot-core/src/test/java/org/apache/pinot/queries/BaseQueriesTest.java:97-102``).  Grammar:

    SELECT item[, item...] FROM table [WHERE cond] [GROUP BY col[, col...]] [ORDER BY expr [ASC|DESC], ...] [LIMIT n]
    item  := COUNT(*) | FN(col) | col            (plain columns are allowed only when they are group-by keys)
    cond  := cond AND cond | cond OR cond | NOT cond | (cond) | col op literal | col BETWEEN a AND b
             | col [NOT] IN (literal, ...)
    op    := = | != | <> | < | <= | > | >=
"""
from __future__ import annotations

import re
from typing import List, Optional

from .query import Aggregation, Filter, FilterNode, Predicate, QueryContext, optimize_filter

_TOKEN = re.compile(r"\s*(?:(\d+\.\d+(?:[eE][-+]?\d+)?|-?\d+\.\d+|-?\d+)|'((?:[^']|'')*)'|(<=|>=|<>|!=|[=<>(),*])|"
                    r"([A-Za-z_$][A-Za-z0-9_$]*))")

_FUNCTIONS = {"COUNT", "SUM", "MIN", "MAX", "AVG", "DISTINCTCOUNT"}


class SqlError(ValueError):
    pass


class _Parser:
    def __init__(self, text: str):
        self.toks = []
        pos = 0
        text = text.strip().rstrip(";")
        while pos < len(text):
            m = _TOKEN.match(text, pos)
            if not m or m.end() == pos:
                raise SqlError(f"cannot tokenize at: {text[pos:pos + 20]!r}")
            num, s, op, ident = m.groups()
            if num is not None:
                self.toks.append(("num", float(num) if ("." in num or "e" in num.lower()) else int(num)))
            elif s is not None:
                self.toks.append(("str", s.replace("''", "'")))
            elif op is not None:
                self.toks.append(("op", op))
            else:
                self.toks.append(("id", ident))
            pos = m.end()
        self.i = 0

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else (None, None)

    def kw(self, word) -> bool:
        k, v = self.peek()
        return k == "id" and v.upper() == word

    def take_kw(self, word) -> bool:
        if self.kw(word):
            self.i += 1
            return True
        return False

    def expect_kw(self, word):
        if not self.take_kw(word):
            raise SqlError(f"expected {word}, got {self.peek()}")

    def take_op(self, op) -> bool:
        k, v = self.peek()
        if k == "op" and v == op:
            self.i += 1
            return True
        return False

    def expect_op(self, op):
        if not self.take_op(op):
            raise SqlError(f"expected {op!r}, got {self.peek()}")

    def ident(self) -> str:
        k, v = self.peek()
        if k != "id":
            raise SqlError(f"expected identifier, got {self.peek()}")
        self.i += 1
        return v

    def literal(self):
        k, v = self.peek()
        if k not in ("num", "str"):
            raise SqlError(f"expected literal, got {self.peek()}")
        self.i += 1
        return v

    # cond := or
    def cond(self) -> FilterNode:
        left = self.and_()
        kids = [left]
        while self.take_kw("OR"):
            kids.append(self.and_())
        return kids[0] if len(kids) == 1 else Filter("OR", kids)

    def and_(self) -> FilterNode:
        kids = [self.not_()]
        while self.kw("AND"):
            self.i += 1
            kids.append(self.not_())
        return kids[0] if len(kids) == 1 else Filter("AND", kids)

    def not_(self) -> FilterNode:
        if self.take_kw("NOT"):
            return Filter("NOT", [self.not_()])
        if self.take_op("("):
            c = self.cond()
            self.expect_op(")")
            return c
        return self.comparison()

    def comparison(self) -> FilterNode:
        col = self.ident()
        if self.take_kw("BETWEEN"):
            lo = self.literal()
            self.expect_kw("AND")
            hi = self.literal()
            return Predicate("RANGE", col, [], lo, hi, True, True)
        negate = self.take_kw("NOT")
        if self.take_kw("IN"):
            self.expect_op("(")
            vals = [self.literal()]
            while self.take_op(","):
                vals.append(self.literal())
            self.expect_op(")")
            return Predicate("NOT_IN" if negate else "IN", col, vals)
        if negate:
            raise SqlError("NOT must be followed by IN here")
        k, op = self.peek()
        if k != "op":
            raise SqlError(f"expected comparison operator, got {self.peek()}")
        self.i += 1
        v = self.literal()
        if op == "=":
            return Predicate("EQ", col, [v])
        if op in ("!=", "<>"):
            return Predicate("NEQ", col, [v])
        if op == ">":
            return Predicate("RANGE", col, [], v, None, False, True)
        if op == ">=":
            return Predicate("RANGE", col, [], v, None, True, True)
        if op == "<":
            return Predicate("RANGE", col, [], None, v, True, False)
        if op == "<=":
            return Predicate("RANGE", col, [], None, v, True, True)
        raise SqlError(f"unsupported operator {op}")


def parse(sql: str, **options) -> QueryContext:
    """SQL text -> QueryContext (filter already run through the broker-side filter optimizers)."""
    p = _Parser(sql)
    p.expect_kw("SELECT")
    aggs: List[Aggregation] = []
    plain: List[str] = []
    while True:
        name = p.ident()
        if name.upper() in _FUNCTIONS and p.take_op("("):
            fn = name.upper()
            if p.take_op("*"):
                if fn != "COUNT":
                    raise SqlError(f"{fn}(*) is not valid")
                aggs.append(Aggregation("COUNT", None))
            else:
                col = p.ident()
                aggs.append(Aggregation("COUNT", None) if fn == "COUNT" else Aggregation(fn, col))
            p.expect_op(")")
            if p.take_kw("FILTER"):   # SUM(x) FILTER (WHERE cond)
                p.expect_op("(")
                p.expect_kw("WHERE")
                aggs[-1].filter = optimize_filter(p.cond())
                p.expect_op(")")
        else:
            plain.append(name)
        if not p.take_op(","):
            break
    p.expect_kw("FROM")
    table = p.ident()
    flt: Optional[FilterNode] = None
    if p.take_kw("WHERE"):
        flt = p.cond()
    group_by: List[str] = []
    if p.take_kw("GROUP"):
        p.expect_kw("BY")
        group_by.append(p.ident())
        while p.take_op(","):
            group_by.append(p.ident())
    order_by = []
    if p.take_kw("ORDER"):
        p.expect_kw("BY")
        while True:
            name = p.ident()
            expr = name
            if p.take_op("("):
                inner = "*" if p.take_op("*") else p.ident()
                p.expect_op(")")
                expr = f"{name.lower()}({inner})"
            asc = True
            if p.take_kw("DESC"):
                asc = False
            else:
                p.take_kw("ASC")
            order_by.append((expr, asc))
            if not p.take_op(","):
                break
    limit = 10
    if p.take_kw("LIMIT"):
        limit = int(p.literal())
    if p.peek()[0] is not None:
        raise SqlError(f"trailing tokens: {p.toks[p.i:]}")
    for c in plain:
        if c not in group_by:
            raise SqlError(f"column {c} must appear in GROUP BY (selection queries are outside this path)")
    if not aggs:
        raise SqlError("only aggregation / group-by queries are on this path")
    return QueryContext(aggregations=aggs, filter=optimize_filter(flt), group_by=group_by, table=table, limit=limit,
                        order_by=order_by, **options)
