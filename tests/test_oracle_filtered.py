"""CPU: the oracle's FILTER (WHERE ...) aggregations (oracle/pinot_oracle.cpp run_query: aggregation infos sharing one group key
generator = AggregationFunctionUtils.buildFilteredAggregationInfos :312-403 + FilteredGroupByOperator.getNextBlock :113-160).

Pinned on the reference's own test: FilteredAggregationsTest.java builds MyTable with INT_COL = i, NO_INDEX_COL = i (raw),
STATIC_INT_COL = 10 for i in [0, 30000) (:77-134) and asserts, query by query, that the FILTER form returns what the
equivalent WHERE / CASE form returns (:146-330).  The same queries (the ones without transform functions) are run through the
oracle here and checked against the CASE semantics computed directly with numpy."""
import numpy as np
import pytest

from pinot_b200 import sql

N = 30_000


@pytest.fixture(scope="module")
def table(oracle):
    i = np.arange(N, dtype=np.int32)
    seg = oracle.build_segment("MyTable", {"INT_COL": i, "NO_INDEX_COL": i.copy(), "STATIC_INT_COL": np.full(N, 10, dtype=np.int32),
                                           "MOD4": (i % 4).astype(np.int32)}, raw=["NO_INDEX_COL"])
    return seg, i.astype(np.int64)


def _agg(fn, vals):
    if fn == "SUM":
        return float(vals.sum())
    if fn == "COUNT":
        return int(len(vals))
    if fn == "MIN":
        return float(vals.min()) if len(vals) else float("inf")
    if fn == "MAX":
        return float(vals.max()) if len(vals) else float("-inf")
    if fn == "AVG":
        return (float(vals.sum()), int(len(vals)))
    raise ValueError(fn)


# (query, [(function, mask over i), ...]) -- aggregation-only cases of testSimpleQueries / testFilterVsCase
CASES = [
    ("SELECT SUM(INT_COL) FILTER(WHERE INT_COL > 9999) FROM MyTable WHERE INT_COL < 1000000", [("SUM", lambda i: i > 9999)]),
    ("SELECT SUM(INT_COL) FILTER(WHERE INT_COL < 3) FROM MyTable WHERE INT_COL > 1", [("SUM", lambda i: (i > 1) & (i < 3))]),
    ("SELECT COUNT(*) FILTER(WHERE INT_COL = 4) FROM MyTable", [("COUNT", lambda i: i == 4)]),
    ("SELECT SUM(INT_COL) FILTER(WHERE INT_COL > 8000) FROM MyTable", [("SUM", lambda i: i > 8000)]),
    ("SELECT SUM(INT_COL) FILTER(WHERE NO_INDEX_COL <= 1) FROM MyTable WHERE INT_COL > 1", [("SUM", lambda i: (i <= 1) & (i > 1))]),
    ("SELECT AVG(INT_COL) FILTER(WHERE NO_INDEX_COL > -1) FROM MyTable", [("AVG", lambda i: i > -1)]),
    ("SELECT MIN(INT_COL) FILTER(WHERE NO_INDEX_COL > 29990), MAX(INT_COL) FILTER(WHERE INT_COL > 29990) FROM MyTable",
     [("MIN", lambda i: i > 29990), ("MAX", lambda i: i > 29990)]),
    ("SELECT SUM(INT_COL) FILTER(WHERE INT_COL > 1234 AND INT_COL < 22000) FROM MyTable", [("SUM", lambda i: (i > 1234) & (i < 22000))]),
    ("SELECT SUM(INT_COL) FILTER(WHERE INT_COL > 3), SUM(INT_COL) FILTER(WHERE INT_COL < 4) FROM MyTable WHERE INT_COL > 2",
     [("SUM", lambda i: (i > 2) & (i > 3)), ("SUM", lambda i: (i > 2) & (i < 4))]),
    ("SELECT SUM(INT_COL) FILTER(WHERE INT_COL > 12345), SUM(INT_COL) FILTER(WHERE INT_COL < 59999), "
     "MIN(INT_COL) FILTER(WHERE INT_COL > 5000) FROM MyTable WHERE INT_COL > 1000",
     [("SUM", lambda i: (i > 1000) & (i > 12345)), ("SUM", lambda i: (i > 1000) & (i < 59999)), ("MIN", lambda i: (i > 1000) & (i > 5000))]),
    ("SELECT SUM(INT_COL) FILTER(WHERE NO_INDEX_COL > 12345), SUM(INT_COL) FILTER(WHERE NO_INDEX_COL < 59999), "
     "MIN(INT_COL) FILTER(WHERE NO_INDEX_COL > 5000) FROM MyTable WHERE INT_COL > 1000",
     [("SUM", lambda i: (i > 1000) & (i > 12345)), ("SUM", lambda i: (i > 1000) & (i < 59999)), ("MIN", lambda i: (i > 1000) & (i > 5000))]),
    ("SELECT SUM(INT_COL) FILTER(WHERE INT_COL > 12345), SUM(NO_INDEX_COL) FILTER(WHERE INT_COL < 59999), "
     "MIN(INT_COL) FILTER(WHERE INT_COL > 5000) FROM MyTable WHERE INT_COL < 28000 AND NO_INDEX_COL > 3000",
     [("SUM", lambda i: (i < 28000) & (i > 3000) & (i > 12345)), ("SUM", lambda i: (i < 28000) & (i > 3000) & (i < 59999)),
      ("MIN", lambda i: (i < 28000) & (i > 3000) & (i > 5000))]),
    ("SELECT SUM(INT_COL) FILTER(WHERE INT_COL > 123 AND INT_COL < 25000), MAX(INT_COL) FILTER(WHERE INT_COL > 123 AND INT_COL < 25000) "
     "FROM MyTable WHERE NO_INDEX_COL > 5 AND NO_INDEX_COL < 29999",
     [("SUM", lambda i: (i > 5) & (i < 29999) & (i > 123) & (i < 25000)), ("MAX", lambda i: (i > 5) & (i < 29999) & (i > 123) & (i < 25000))]),
    # a FILTER next to non-filtered functions (the main info)
    ("SELECT SUM(INT_COL) FILTER(WHERE INT_COL > 20000), SUM(NO_INDEX_COL), MAX(INT_COL), COUNT(*) FROM MyTable WHERE NO_INDEX_COL > 5",
     [("SUM", lambda i: (i > 5) & (i > 20000)), ("SUM", lambda i: i > 5), ("MAX", lambda i: i > 5), ("COUNT", lambda i: i > 5)]),
]


@pytest.mark.parametrize("text,expected", CASES, ids=[str(k) for k in range(len(CASES))])
def test_reference_filtered_aggregation_queries(oracle, table, text, expected):
    seg, i = table
    r = oracle.execute(seg, sql.parse(text))
    assert r.num_groups == -1
    for a, (fn, mask) in enumerate(expected):
        want = _agg(fn, i[mask(i)])
        got = (float(r.doubles[a][0]), int(r.longs[a][0])) if fn == "AVG" else int(r.longs[a][0]) if fn == "COUNT" else float(r.doubles[a][0])
        assert got == want, (text, a, fn, got, want)


def test_group_by_keeps_every_group_of_the_main_filter(oracle, table):
    """FilteredGroupByOperator: the main-filter info exists even without a function in it, so groups no FILTER clause matches
    still appear, with the functions' defaults (AggregationFunctionUtils.java:388-400)."""
    seg, i = table
    q = sql.parse("SELECT SUM(INT_COL) FILTER(WHERE INT_COL > 9999), COUNT(*) FILTER(WHERE MOD4 = 1), MIN(INT_COL) FILTER(WHERE MOD4 = 1) "
                  "FROM MyTable WHERE INT_COL < 1000000 GROUP BY MOD4")
    r = oracle.execute(seg, q)
    assert r.num_groups == 4
    for g in range(4):
        k = seg.value_of("MOD4", int(r.keys[g, 0]))
        rows = i[i % 4 == k]
        assert float(r.doubles[0][g]) == float(rows[rows > 9999].sum())
        assert int(r.longs[1][g]) == (len(rows) if k == 1 else 0)
        assert float(r.doubles[2][g]) == (float(rows.min()) if k == 1 else float("inf"))
    # statistics add up over the infos (FilteredGroupByOperator.java:148-150): 3 infos (2 clauses + main)
    assert r.stats[0] == int((i > 9999).sum()) + int((i % 4 == 1).sum()) + N


def test_filter_equals_where_for_group_by(oracle, table):
    """testFilterResultColumnNameGroupBy: FILTER(c) ... WHERE m GROUP BY k == ... WHERE c AND m GROUP BY k on the groups the
    latter produces; the remaining groups of the main filter hold the default."""
    seg, _ = table
    a = oracle.execute(seg, sql.parse("SELECT SUM(INT_COL) FILTER(WHERE INT_COL > 9999 AND INT_COL < 20000) FROM MyTable GROUP BY MOD4"))
    b = oracle.execute(seg, sql.parse("SELECT SUM(INT_COL) FROM MyTable WHERE INT_COL > 9999 AND INT_COL < 20000 GROUP BY MOD4"))
    ta = {int(a.keys[g, 0]): float(a.doubles[0][g]) for g in range(a.num_groups)}
    tb = {int(b.keys[g, 0]): float(b.doubles[0][g]) for g in range(b.num_groups)}
    assert ta == tb and len(ta) == 4
