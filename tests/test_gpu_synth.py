"""-m gpu: the device-side synthetic segment creator writes Pinot's exact bytes, and full-size (100 M-row) runs obey
size-independent properties."""
import numpy as np
import pytest

from gpu_util import assert_tables_equal, gpu_table, oracle_table
from oracle import segment_builder as sb
from pinot_b200 import sql
from pinot_b200.plan_maker import B200Context, B200PlanMaker, IndexSegment

pytestmark = pytest.mark.gpu

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def synth_dict_ids(seed, n, card):
    """numpy twin of pb200_synth.cu: mix64(seed + doc * golden) % card (SplitMix64 finaliser)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z % np.uint64(card)).astype(np.int32)


@pytest.fixture(scope="module")
def ctx():
    c = B200Context(0)
    yield c
    c.close()


SPECS = [{"name": "a", "cardinality": 10, "seed": 11, "inverted": True},
         {"name": "b", "cardinality": 1000, "seed": 12, "value_base": 5, "value_step": 3, "inverted": True},
         {"name": "c", "cardinality": 100_000, "seed": 13, "value_base": -7, "value_step": 7},
         {"name": "d", "cardinality": 65_536, "seed": 14},
         {"name": "e", "cardinality": 2, "seed": 15, "inverted": True},
         {"name": "f", "cardinality": 1, "seed": 16}]


@pytest.mark.parametrize("n", [1, 100, 65_536, 300_007])
def test_synthetic_segment_bytes_match_the_reference_writers(oracle, ctx, n):
    seg = IndexSegment.synthetic(ctx, "syn", n, SPECS)
    try:
        cols = []
        for spec in SPECS:
            card = spec["cardinality"]
            ids = synth_dict_ids(spec["seed"], n, card)
            bits = sb.num_bits_per_value(card - 1)
            info = seg.column_info(spec["name"])
            assert (info["bits"], info["cardinality"]) == (bits, card)
            fwd = seg.read_index(spec["name"], "fwd")
            assert np.array_equal(fwd, sb.pack_fixed_bits(ids, bits)), spec["name"]           # numpy packer
            if n <= 65_536:
                assert np.array_equal(fwd, oracle.bitset_write(ids, bits)), spec["name"]      # PinotDataBitSet.writeInt
            values = (spec.get("value_base", 0) + spec.get("value_step", 1) * np.arange(card)).astype(np.int32)
            dct = seg.read_index(spec["name"], "dict")
            assert np.array_equal(np.frombuffer(dct.tobytes(), dtype=">i4"), values)           # BE sorted INT dictionary
            inv = None
            if spec.get("inverted"):
                inv = seg.read_index(spec["name"], "inv")
                offs = np.frombuffer(inv[: 4 * (card + 1)].tobytes(), dtype=">u4").astype(np.int64)
                assert offs[0] == 4 * (card + 1) and offs[-1] == len(inv)                      # BitmapInvertedIndexWriter
                for d in range(card) if card <= 10 else [0, 1, card // 2, card - 1]:
                    docs = oracle.roaring_deserialize(inv[offs[d]:offs[d + 1]])
                    assert np.array_equal(docs, np.nonzero(ids == d)[0]), (spec["name"], d)
            cols.append(sb.ColumnData(spec["name"], sb.INT, True, bits, card, False, 4, fwd, dct, inv,
                                      dict_values=values, dict_ids=ids))
        host = sb.SegmentData("syn", n, cols)
        pm = B200PlanMaker(ctx)
        for text in ("SELECT COUNT(*), SUM(c), MIN(d), MAX(d) FROM t WHERE b BETWEEN 100 AND 900 AND c > 100000",
                     "SELECT COUNT(*), SUM(c) FROM t WHERE a = 3 AND b = 302 AND e = 1",           # 3 inverted EQ, AND-ed
                     "SELECT SUM(c), MAX(d) FROM t WHERE a IN (1, 2, 9) AND e != 0 GROUP BY b",
                     "SELECT COUNT(*) FROM t WHERE a = 4 GROUP BY a, e"):
            q = sql.parse(text)
            want = oracle_table(host, q, oracle.execute(host, q))
            got = gpu_table(host, q, pm.make_segment_plan_node(seg, q).run().next_block())
            assert_tables_equal(q, got, want, f"n={n}: {text}")
    finally:
        seg.destroy()


def test_full_size_segment_properties(ctx):
    """BASELINE.json's full size (100 M rows): properties that need no oracle run."""
    n = 100_000_000
    specs = [{"name": "g", "cardinality": 1000, "seed": 1}, {"name": "x", "cardinality": 100_000, "seed": 2,
                                                              "value_base": 3, "value_step": 5},
             {"name": "y", "cardinality": 1_000_000, "seed": 3}]
    seg = IndexSegment.synthetic(ctx, "big", n, specs)
    pm = B200PlanMaker(ctx)
    try:
        run = lambda text: pm.make_segment_plan_node(seg, sql.parse(text)).run().next_block()
        total = run("SELECT COUNT(*), SUM(x), MIN(x), MAX(x), SUM(y) FROM t WHERE g >= 0 AND y >= 0")
        # MATCH_ALL leaves collapse -> every row counted exactly once
        assert int(total.longs[0][0]) == n
        lo = run("SELECT COUNT(*), SUM(x), SUM(y) FROM t WHERE y < 400000")
        hi = run("SELECT COUNT(*), SUM(x), SUM(y) FROM t WHERE y >= 400000")
        assert int(lo.longs[0][0]) + int(hi.longs[0][0]) == n                         # partition: counts add up
        assert lo.doubles[1][0] + hi.doubles[1][0] == total.doubles[1][0]             # ... and exact integer sums
        assert lo.doubles[2][0] + hi.doubles[2][0] == total.doubles[4][0]
        assert abs(int(lo.longs[0][0]) - 0.4 * n) < 5e-4 * n                          # uniform dictIds
        grouped = run("SELECT COUNT(*), SUM(x), MIN(x), MAX(x) FROM t GROUP BY g")
        assert grouped.num_groups == 1000 and int(grouped.longs[0].sum()) == n        # checksum of checksums
        assert grouped.doubles[1].sum() == total.doubles[1][0]
        assert grouped.doubles[2].min() == total.doubles[2][0] and grouped.doubles[3].max() == total.doubles[3][0]
        again = run("SELECT COUNT(*), SUM(x), MIN(x), MAX(x) FROM t GROUP BY g")       # idempotent / deterministic
        assert np.array_equal(again.doubles[1], grouped.doubles[1]) and np.array_equal(again.keys, grouped.keys)
        # first rows reproduce the numpy twin of the generator (spot check of the 100 M-row buffers)
        fwd = seg.read_index("y", "fwd")[: 20 * 4096 // 8]
        assert np.array_equal(sb.unpack_fixed_bits(fwd, 4096, 20), synth_dict_ids(3, 4096, 1_000_000))
    finally:
        seg.destroy()
