"""-m gpu: the native segment-directory loader (pb200h_segment_load_dir = ImmutableSegmentLoader.load without a JVM):
v1 file-per-index and v3 columns.psf + index_map layouts, and StarTreeV2 files picked up from the directory."""
import numpy as np
import pytest

from gpu_util import assert_tables_equal, check_query, gpu_table, oracle_table
from oracle import startree_builder as stb
from pinot_b200 import sql
from pinot_b200.plan_maker import B200Context, B200PlanMaker, IndexSegment
from segment_dir_util import write_segment_dir
from test_gpu_parity import QUERIES, _random_segment

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = B200Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("version", ["v1", "v3"])
def test_load_segment_directory(oracle, ctx, tmp_path, version):
    rng = np.random.default_rng(11)
    seg = _random_segment(oracle, rng, 20_011, name="disk")
    root = write_segment_dir(str(tmp_path / version), seg, version)
    dev = IndexSegment.load(ctx, root)
    pm = B200PlanMaker(ctx)
    try:
        assert dev.num_docs == seg.num_docs and sorted(dev.column_names) == sorted(c.name for c in seg.columns)
        for c in seg.columns:  # the loader reproduces what from_columns gets directly
            info = dev.column_info(c.name)
            assert (info["bits"], info["cardinality"], bool(info["is_sorted"])) == (c.bits, c.cardinality, c.is_sorted), c.name
            assert bool(info["has_inverted"]) == (c.inv is not None and not c.is_sorted), c.name
        for text in QUERIES:
            check_query(oracle, pm, seg, dev, sql.parse(text), what=f"{version}: {text}")
    finally:
        dev.destroy()


@pytest.mark.parametrize("version", ["v1", "v3"])
def test_load_directory_with_compressed_raw_columns(oracle, ctx, tmp_path, version):
    """Raw metric / dimension columns as a table config with noDictionaryColumns writes them (`<col>.sv.raw.fwd` /
    `forward_index` in columns.psf; Snappy, LZ4 and PASS_THROUGH chunks): loaded from disk, decoded and dictionary-encoded
    at load, queried as aggregation arguments, filter columns and GROUP BY keys."""
    rng = np.random.default_rng(13)
    n = 12_345
    seg = oracle.build_segment("raw_disk", {
        "ts": (1_700_000_000_000 + rng.integers(0, 500, size=n).astype(np.int64) * 60_000),     # raw LONG, LZ4 (dimension default)
        "price": (rng.integers(0, 2000, size=n) / 100.0).astype(np.float64),                     # raw DOUBLE, PASS_THROUGH (metric default)
        "qty": rng.integers(1, 50, size=n).astype(np.int32),                                     # raw INT, SNAPPY
        "k": rng.integers(0, 12, size=n).astype(np.int32)},
        raw=["ts", "price", "qty"], raw_compression={"ts": 3, "price": 0, "qty": 1})
    root = write_segment_dir(str(tmp_path / version), seg, version)
    dev = IndexSegment.load(ctx, root)
    pm = B200PlanMaker(ctx)
    try:
        for text in ("SELECT SUM(price), MAX(ts), MIN(qty), COUNT(*) FROM t WHERE ts >= 1700000600000 AND qty < 40",
                     "SELECT SUM(qty), AVG(price) FROM t WHERE price BETWEEN 2.5 AND 17.25 GROUP BY k",
                     "SELECT COUNT(*), SUM(price) FROM t WHERE k != 3 GROUP BY ts",
                     "SELECT MAX(price) FROM t GROUP BY qty, k"):
            check_query(oracle, pm, seg, dev, sql.parse(text, num_groups_limit=1_000_000), what=f"{version}: {text}")
    finally:
        dev.destroy()


def test_load_directory_with_star_tree(oracle, ctx, tmp_path):
    rng = np.random.default_rng(12)
    n = 40_000
    seg = oracle.build_segment("st_disk", {
        "d1": rng.integers(0, 60, size=n).astype(np.int32), "d2": rng.integers(0, 40, size=n).astype(np.int32) * 3,
        "d3": rng.integers(0, 5, size=n).astype(np.int32), "m": rng.integers(0, 500, size=n).astype(np.int32)})
    st = stb.build_star_tree(seg, ["d1", "d2", "d3"], [("COUNT", None), ("SUM", "m"), ("MAX", "m")], max_leaf_records=100)
    root = write_segment_dir(str(tmp_path / "seg"), seg, "v3", star_tree=st)
    dev = IndexSegment.load(ctx, root)
    pm = B200PlanMaker(ctx)
    try:
        for text in ("SELECT COUNT(*), SUM(m), MAX(m), AVG(m) FROM t WHERE d1 < 20 GROUP BY d2",
                     "SELECT SUM(m) FROM t WHERE d1 = 7 AND d3 IN (1, 2)",
                     "SELECT COUNT(*) FROM t GROUP BY d3, d1"):
            q = sql.parse(text)
            star = pm.execute_segments([dev], q)[0]
            assert star.operator_kind == "STAR_TREE", text          # the tree came from the directory
            scan = pm.execute_segments([dev], sql.parse(text, use_star_tree=False))[0]
            want = oracle_table(seg, q, oracle.execute(seg, q))
            assert_tables_equal(q, gpu_table(seg, q, scan), want, "scan " + text)
            assert_tables_equal(q, gpu_table(seg, q, star), want, "star " + text)
            assert star.stats.num_docs_scanned <= scan.stats.num_docs_scanned
    finally:
        dev.destroy()


def test_load_reference_written_v1_directory(oracle, ctx):
    """tests/golden/paddingOld/: a segment directory WRITTEN BY THE REFERENCE (pinot-core/src/test/resources/data/
    paddingOld.tar.gz, v1 layout with its own metadata.properties; extracted by tests/golden/make_golden.py).  It predates
    segment.padding.character, so its STRING column is padded with '%': like the reference (ColumnMetadataImpl.java:250-253)
    only zero padding is accepted -- the loader leaves `name` out and loads the INT / FLOAT / LONG columns, whose decoded
    contents and query results are checked against the golden bytes."""
    import os
    from oracle import segment_builder as sb
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "paddingOld")
    dev = IndexSegment.load(ctx, root)
    pm = B200PlanMaker(ctx)
    try:
        assert dev.num_docs == 5
        assert sorted(dev.column_names) == ["age", "outgoingName1", "percent"]     # `name`: legacy '%' padding, not loaded
        cols = []
        for name, dt, be in (("age", sb.INT, ">i4"), ("percent", sb.FLOAT, ">f4"), ("outgoingName1", sb.LONG, ">i8")):
            dct = np.frombuffer(open(os.path.join(root, name + ".dict"), "rb").read(), dtype=np.uint8)
            fwd = np.frombuffer(open(os.path.join(root, name + ".sv.unsorted.fwd"), "rb").read(), dtype=np.uint8)
            vals = np.frombuffer(dct.tobytes(), dtype=be)
            info = dev.column_info(name)
            assert (info["bits"], info["cardinality"]) == (3, 5), name
            assert [dev.dictionary_value(name, i) for i in range(5)] == [v.item() for v in vals.astype(be[1:])], name
            assert np.array_equal(dev.read_index(name, "fwd"), fwd), name                     # the file's bytes, back from HBM
            cols.append(sb.ColumnData(name, dt, True, 3, 5, False, vals.dtype.itemsize, fwd, dct, None,
                                      dict_values=vals.astype(be[1:]), dict_ids=sb.unpack_fixed_bits(fwd, 5, 3)))
        seg = sb.SegmentData("mySegment_0", 5, cols)
        assert [seg.value_of("age", int(i)) for i in cols[0].dict_ids] == [1228, 837, 1209, 617, 824]
        for text in ("SELECT COUNT(*), SUM(age), MIN(percent), MAX(outgoingName1), AVG(age) FROM t",
                     "SELECT COUNT(*), SUM(outgoingName1) FROM t WHERE age > 800 AND percent < 900.0",
                     "SELECT SUM(percent), MAX(age) FROM t WHERE outgoingName1 >= 310 GROUP BY age",
                     "SELECT DISTINCTCOUNT(age), COUNT(*) FROM t WHERE age IN (617, 1228, 5)"):
            check_query(oracle, pm, seg, dev, sql.parse(text), what="paddingOld: " + text)
    finally:
        dev.destroy()


def test_segment_cache_residency(oracle, ctx, tmp_path):
    """pb200h_cache_*: keyed by (name, CRC), pinned while a query holds the segment, LRU eviction under a byte budget,
    refresh (same name, new CRC) and drop while pinned are deferred to the last release."""
    from pinot_b200._lib import Pb200Error
    from pinot_b200.plan_maker import SegmentCache
    rng = np.random.default_rng(31)
    segs, roots = [], []
    for i in range(4):
        s = oracle.build_segment(f"cs{i}", {"a": rng.integers(0, 50, size=60_000).astype(np.int32),
                                            "b": rng.integers(0, 5000, size=60_000).astype(np.int32) + i})
        segs.append(s)
        roots.append(write_segment_dir(str(tmp_path / f"cs{i}"), s, "v1"))
    probe = IndexSegment.load(ctx, roots[0])
    one = probe.device_bytes()
    probe.destroy()
    cache = SegmentCache(ctx, max_device_bytes=int(2.5 * one))       # room for two segments
    pm = B200PlanMaker(ctx)
    q = sql.parse("SELECT COUNT(*), SUM(b) FROM t WHERE a < 25 GROUP BY a")
    try:
        s0 = cache.acquire("cs0", 100, roots[0], one)
        check_query(oracle, pm, segs[0], s0, q, "cached cs0")
        s1 = cache.acquire("cs1", 101, roots[1], one)
        with pytest.raises(Pb200Error):                                # both residents are pinned: no room for a third
            cache.acquire("cs2", 102, roots[2], one)
        cache.release(s1)
        s2 = cache.acquire("cs2", 102, roots[2], one)                 # evicts the unpinned cs1, never the pinned cs0
        st = cache.stats()
        assert (st["segments"], st["evictions"], st["misses"]) == (2, 1, 4) and st["bytes"] <= st["budget"]
        again = cache.acquire("cs0", 100)                             # hit: no directory needed
        assert cache.stats()["hits"] == 1
        cache.release(again)
        cache.release(s0)
        cache.release(s2)
        # LRU: cs0 was touched last, so loading cs3 evicts cs2
        s3 = cache.acquire("cs3", 103, roots[3], one)
        cache.acquire("cs0", 100)
        with pytest.raises(Pb200Error):
            cache.acquire("cs2", 102)                                  # gone, and no directory given
        # refresh: same name, new CRC -> the old version stays usable for its holder and is freed by its release
        cache.release(s3)
        new0 = cache.acquire("cs0", 999, roots[1], one)               # content of cs1 under the name cs0
        check_query(oracle, pm, segs[1], new0, q, "refreshed cs0")
        st = cache.stats()
        assert st["segments"] >= 2
        cache.evict("cs0", 999)                                        # dropped while pinned: freed by the release
        check_query(oracle, pm, segs[1], new0, q, "dropped but still held")
        cache.release(new0)
    finally:
        cache.close()
