"""CPU: the merge / final-result helpers of pinot_b200.cli (what the reference's combine + broker reduce do above the path)
against tests/reduce_util (the test-side restatement used by the golden inter-segment tests)."""
from types import SimpleNamespace

import numpy as np

from pinot_b200 import cli, sql
from reduce_util import combine, normalise, reduce_rows


class _Seg:
    """dictionary_value over an oracle-built segment (stands in for a loaded IndexSegment)."""

    def __init__(self, data):
        self.data = data

    def dictionary_value(self, column, dict_id):
        return self.data.value_of(column, dict_id)


def test_cli_merge_equals_reduce_util(oracle):
    rng = np.random.default_rng(3)
    segs = [oracle.build_segment(f"s{i}", {"k": rng.integers(0, 5 + i, size=4000).astype(np.int32) * 3,
                                            "v": rng.integers(-50, 50, size=4000).astype(np.int32)}) for i in range(3)]
    for text in ("SELECT COUNT(*), SUM(v), MIN(v), MAX(v), AVG(v), DISTINCTCOUNT(v) FROM t WHERE v > -40 GROUP BY k LIMIT 100",
                 "SELECT COUNT(*), AVG(v), DISTINCTCOUNT(k) FROM t"):
        q = sql.parse(text)
        results = [oracle.execute(s, q) for s in segs]
        blocks = [SimpleNamespace(num_groups=r.num_groups, keys=r.keys, doubles=r.doubles, longs=r.longs, distinct=r.distinct) for r in results]
        got = cli.merge_blocks(q, [_Seg(s) for s in segs], blocks)
        want = combine([a.function for a in q.aggregations], [normalise(s, q, r.num_groups, r.keys, r.doubles, r.longs, r.distinct)
                                                                for s, r in zip(segs, results)])
        assert set(got) == set(want)
        for key in want:
            for a, agg in enumerate(q.aggregations):
                g, w = got[key][a], want[key][a]
                assert (frozenset(g) == w) if agg.function == "DISTINCTCOUNT" else (g == w), (text, key, agg.function)
        rows = cli.final_rows(q, got)
        ref = {k: v for k, v in reduce_rows([a.function for a in q.aggregations], want)}
        assert len(rows) == min(q.limit, len(want))
        for r in rows:
            key = tuple(r[: len(q.group_by)])
            assert r[len(q.group_by):] == ref[key]
