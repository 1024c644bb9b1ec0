"""-m gpu: results as DataTableImplV4 bytes (pb200h_result_to_datatable), read back with the independent test-side
reader (tests/datatable_util.py) and compared, in VALUE space, with the CPU oracle on the reference's own sv fixture and
on a segment with every dictionary type."""
import numpy as np
import pytest

import datatable_util
from gpu_util import oracle_table, to_device, SUM_REL_TOL
from pinot_b200 import sql
from pinot_b200.plan_maker import B200Context, B200PlanMaker

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = B200Context(0)
    yield c
    c.close()


def _check(oracle, ctx, seg, dev, text, merge_copies=1):
    pm = B200PlanMaker(ctx)
    q = sql.parse(text, num_groups_limit=1_000_000)
    block = pm.execute_segments([dev] * merge_copies, q, merge=merge_copies > 1, keep_handle=True, defer=False)[0]
    try:
        dt = datatable_util.parse(pm.to_datatable(dev, q, block))
    finally:
        block.release(ctx)
    want = oracle_table(seg, q, oracle.execute(seg, q))
    ngb = len(q.group_by)
    assert dt["column_names"] == q.group_by + [f"{a.function.lower()}({a.column or '*'})" for a in q.aggregations]
    expect_types = [{0: "INT", 1: "LONG", 2: "FLOAT", 3: "DOUBLE", 4: "STRING"}[seg.column(c).data_type] for c in q.group_by] + \
                   [{"COUNT": "LONG", "AVG": "OBJECT", "DISTINCTCOUNT": "OBJECT"}.get(a.function, "DOUBLE") for a in q.aggregations]
    assert dt["column_types"] == expect_types
    assert len(dt["rows"]) == len(want)
    got = {tuple(r[:ngb]): r[ngb:] for r in dt["rows"]}
    for key, wv in want.items():
        k = tuple(np.float32(x).item() if isinstance(x, float) and seg.column(c).data_type == 2 else x for x, c in zip(key, q.group_by))
        gv = got[k]
        for a, agg in enumerate(q.aggregations):
            w = wv[a]
            if agg.function in ("COUNT",):
                assert gv[a] == w * merge_copies
            elif agg.function == "SUM":
                assert abs(gv[a] - w * merge_copies) <= SUM_REL_TOL * abs(w * merge_copies) + 1e-9
            elif agg.function == "AVG":
                assert gv[a][1] == w[1] * merge_copies and abs(gv[a][0] - w[0] * merge_copies) <= SUM_REL_TOL * abs(w[0] * merge_copies) + 1e-9
            elif agg.function == "DISTINCTCOUNT":
                assert gv[a] == w, (text, key)
            else:
                assert gv[a] == w, (text, key, agg.function, gv[a], w)
    stats = oracle.execute(seg, q).stats
    assert dt["metadata"]["numDocsScanned"] == stats[0] * merge_copies and dt["metadata"]["totalDocs"] == stats[3] * merge_copies
    assert dt["metadata"]["numGroupsLimitReached"] == "false"
    return dt


def test_datatable_of_the_reference_fixture(oracle, ctx, sv_segment):
    dev = to_device(ctx, sv_segment)
    try:
        for text in ("SELECT COUNT(*), SUM(column1), MAX(column3), MIN(column6), AVG(column7) FROM testTable",
                     "SELECT COUNT(*), SUM(column1), MAX(column3), AVG(column6) FROM testTable WHERE column1 > 100000000 GROUP BY column11",   # STRING key
                     "SELECT SUM(column3), DISTINCTCOUNT(column9) FROM testTable GROUP BY column11, column6",
                     "SELECT DISTINCTCOUNT(column11), DISTINCTCOUNT(column1) FROM testTable WHERE column3 < 1000000000"):
            dt = _check(oracle, ctx, sv_segment, dev, text)
            assert dt["version"] == 4
        _check(oracle, ctx, sv_segment, dev, "SELECT COUNT(*), SUM(column1), AVG(column6) FROM testTable GROUP BY column11", merge_copies=3)
    finally:
        dev.destroy()


def test_datatable_key_types(oracle, ctx):
    rng = np.random.default_rng(8)
    n = 30_000
    seg = oracle.build_segment("dt", {
        "i": rng.integers(-50, 50, size=n).astype(np.int32), "l": rng.integers(0, 40, size=n).astype(np.int64) * 10_000_000_007,
        "f": rng.integers(0, 30, size=n).astype(np.float32) * 0.25, "d": rng.integers(0, 30, size=n).astype(np.float64) / 7.0,
        "s": np.array([b"a", b"bb", b"ccc", b"dddd e"])[rng.integers(0, 4, size=n)], "v": rng.integers(0, 1000, size=n).astype(np.int32)})
    dev = to_device(ctx, seg)
    try:
        for text in ("SELECT SUM(v), COUNT(*) FROM t GROUP BY i, s", "SELECT MAX(v), MIN(d), AVG(v) FROM t WHERE v > 10 GROUP BY l, f",
                     "SELECT COUNT(*), DISTINCTCOUNT(l), DISTINCTCOUNT(f), DISTINCTCOUNT(d), DISTINCTCOUNT(s) FROM t GROUP BY d",
                     "SELECT COUNT(*) FROM t WHERE v > 5000 GROUP BY s"):   # no rows
            _check(oracle, ctx, seg, dev, text)
    finally:
        dev.destroy()
