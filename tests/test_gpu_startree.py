"""-m gpu: star-tree execution on the device == plain scan on the device == the oracle (BaseStarTreeV2Test property)."""
import numpy as np
import pytest

from gpu_util import assert_tables_equal, gpu_table, oracle_table, to_device
from oracle import startree_builder as stb
from oracle.startree_query import execute_with_star_tree
from pinot_b200 import sql
from pinot_b200.plan_maker import B200Context, B200PlanMaker
from test_oracle_startree import FILTERS, GROUPS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = B200Context(0)
    yield c
    c.close()


def _attach(dev, st):
    nd = len(st.dimensions)
    dims = [st.segment.columns[j].fwd for j in range(nd)]
    metrics = [(fn, col, st.segment.columns[nd + i].fwd) for i, (fn, col) in enumerate(st.function_pairs)]
    dev.attach_star_tree(st.tree, st.num_docs, st.dimensions, dims, metrics)


@pytest.mark.parametrize("max_leaf", [1, 10, 1000])
def test_star_tree_on_device_equals_scan(oracle, ctx, max_leaf):
    rng = np.random.default_rng(max_leaf)
    n = 50_000
    seg = oracle.build_segment("st", {
        "d1": rng.integers(0, 100, size=n).astype(np.int32),
        "d2": rng.integers(0, 100, size=n).astype(np.int32) * 2,
        "d3": rng.integers(0, 7, size=n).astype(np.int32),
        "m": rng.integers(0, 1000, size=n).astype(np.int32),
    })
    st = stb.build_star_tree(seg, ["d1", "d2", "d3"], [("COUNT", None), ("SUM", "m"), ("MAX", "m"), ("MIN", "m")],
                             max_leaf_records=max_leaf)
    dev = to_device(ctx, seg)
    pm = B200PlanMaker(ctx)
    try:
        _attach(dev, st)
        for flt in FILTERS:
            for gb in GROUPS:
                text = "SELECT COUNT(*), SUM(m), MAX(m), MIN(m), AVG(m) FROM t" + flt + gb
                q = sql.parse(text)
                star = pm.execute_segments([dev], q)[0]
                assert star.operator_kind in ("STAR_TREE", "EMPTY", "NON_SCAN_AGGREGATION"), text
                q_scan = sql.parse(text, use_star_tree=False)
                scan = pm.execute_segments([dev], q_scan)[0]
                assert scan.operator_kind != "STAR_TREE"
                want = oracle_table(seg, q, oracle.execute(seg, q))
                got_scan = gpu_table(seg, q, scan)
                assert_tables_equal(q, got_scan, want, "scan " + text)
                if star.operator_kind == "STAR_TREE":
                    got_star = gpu_table(seg, q, star)
                    if not gb and scan.stats.num_docs_scanned == 0:
                        continue  # empty aggregation-only result: MIN/MAX defaults are representation only
                    assert_tables_equal(q, got_star, want, "star " + text)
                    assert got_star == execute_with_star_tree(oracle, seg, st, q), text
                    if not flt:
                        assert star.stats.num_docs_scanned < scan.stats.num_docs_scanned  # pre-aggregation pays off
        # a query that does not fit must fall back to the scan operators
        b = pm.execute_segments([dev], sql.parse("SELECT SUM(m) FROM t WHERE d1 = 1 OR d2 = 2"))[0]
        assert b.operator_kind == "AGGREGATION"
        b = pm.execute_segments([dev], sql.parse("SELECT DISTINCTCOUNT(m) FROM t WHERE d1 = 1"))[0]
        assert b.operator_kind == "AGGREGATION"
    finally:
        dev.destroy()
