"""Pins the CPU oracle to the reference's own known-answer tests (no GPU needed).

Source of every expected number: golden_cases.py (transcribed from InnerSegmentAggregationSingleValueQueriesTest,
InterSegmentAggregationSingleValueQueriesTest, InterSegmentGroupBySingleValueQueriesTest with file:line).
"""
import numpy as np
import pytest

import golden_cases as G
from pinot_b200 import sql
from reduce_util import combine, normalise, reduce_rows


def _norm(seg, q, r):
    return normalise(seg, q, r.num_groups, r.keys, r.doubles, r.longs, r.distinct)


@pytest.mark.parametrize("query,stats,expected", G.INNER_SEGMENT_AGGREGATION)
def test_inner_segment_aggregation(oracle, sv_segment, query, stats, expected):
    q = sql.parse(query)
    r = oracle.execute(sv_segment, q)
    assert r.stats == stats
    count, s, mx, mn, avg_sum, avg_cnt = expected
    assert int(r.longs[0][0]) == count
    assert r.doubles[1][0] == float(s)
    assert r.doubles[2][0] == float(mx)
    assert r.doubles[3][0] == float(mn)
    assert r.doubles[4][0] == float(avg_sum) and int(r.longs[4][0]) == avg_cnt


@pytest.mark.parametrize("query,regime,stats,key,expected", G.INNER_SEGMENT_GROUP_BY)
def test_inner_segment_group_by(oracle, sv_segment, query, regime, stats, key, expected):
    q = sql.parse(query)
    r = oracle.execute(sv_segment, q)
    assert r.regime == regime
    assert r.stats == stats
    table = _norm(sv_segment, q, r)
    count, s, mx, mn, avg_sum, avg_cnt = expected
    assert table[tuple(key)] == [count, float(s), float(mx), float(mn), (float(avg_sum), avg_cnt)]


def _four_segments(oracle, seg, query):
    q = sql.parse(query)
    fns = [a.function for a in q.aggregations]
    blocks = [_norm(seg, q, oracle.execute(seg, q)) for _ in range(4)]
    return q, fns, reduce_rows(fns, combine(fns, blocks))


@pytest.mark.parametrize("select,ascending,expected", G.INTER_SEGMENT)
def test_inter_segment(oracle, sv_segment, select, ascending, expected):
    base = f"SELECT {select} FROM testTable"
    for i, (flt, gb) in enumerate([("", ""), (G.FILTER, ""), ("", G.INTER_GROUP_BY), (G.FILTER, G.INTER_GROUP_BY)]):
        _, _, rows = _four_segments(oracle, sv_segment, base + flt + gb)
        if gb:
            sign = 1 if ascending else -1
            rows.sort(key=lambda kv: (sign * kv[1][0], sign * kv[1][1]))  # ORDER BY v1, v2 [DESC] LIMIT 1
        got = rows[0][1]
        for g, e in zip(got, expected[i]):
            assert g == pytest.approx(e, rel=1e-9), (select, i)


def test_inter_segment_count(oracle, sv_segment):
    got = []
    for flt, gb in [("", ""), (G.FILTER, ""), ("", G.INTER_GROUP_BY), (G.FILTER, G.INTER_GROUP_BY)]:
        _, _, rows = _four_segments(oracle, sv_segment, "SELECT COUNT(*) FROM testTable" + flt + gb)
        rows.sort(key=lambda kv: -kv[1][0])
        got.append(rows[0][1][0])
    assert got == G.INTER_SEGMENT_COUNT
    r = oracle.execute(sv_segment, sql.parse("SELECT COUNT(*) FROM testTable" + G.FILTER))
    assert (4 * r.stats[0], 4 * r.stats[1]) == G.INTER_SEGMENT_FILTER_STATS


def test_inter_segment_group_by_string_key(oracle, sv_segment):
    _, _, rows = _four_segments(oracle, sv_segment, "SELECT column11, SUM(column1) FROM testTable GROUP BY column11")
    rows.sort(key=lambda kv: kv[0])
    assert [(k[0], v[0]) for k, v in rows] == G.INTER_GROUP_BY_COLUMN11_SUM


def test_brute_force_cross_check(oracle, sv_segment, sv_columns):
    """The goldens again, from the raw values with numpy (guards the fixture <-> segment builder chain)."""
    c = sv_columns
    m = ((c["column1"] > 100000000) & (c["column3"] >= 20000000) & (c["column3"] <= 1000000000) &
         (c["column5"] == b"gFuH") & ((c["column6"] < 500000000) | ~np.isin(c["column11"], [b"t", b"P"])) &
         (c["daysSinceEpoch"] == 126164076))
    assert int(m.sum()) == 6129
    assert int(c["column1"][m].astype(np.int64).sum()) == 6875947596072
    ids, _ = oracle.filter_doc_ids(sv_segment, sql.parse("SELECT COUNT(*) FROM testTable" + G.FILTER))
    assert np.array_equal(ids, np.nonzero(m)[0])
