"""CPU: the oracle's restatement of the reference's no-dictionary paths.

  * chunk codecs (oracle/chunk_codecs.py) against the reference-written raw forward-index files (golden/raw_forward/,
    FixedByteChunkSVForwardIndexTest.java:340-377);
  * NoDictionarySingleColumnGroupKeyGenerator.java:60-147 / NoDictionaryMultiColumnGroupKeyGenerator.java as chosen by
    DefaultGroupByExecutor.java:87-117: grouping by a raw column must give, value for value, what grouping by the same data
    dictionary-encoded gives (the reference's tests check exactly this equivalence against a hash map:
    NoDictionaryGroupKeyGeneratorTest.java), group ids are handed out in first-seen order and numGroupsLimit cuts there.
"""
import os

import numpy as np
import pytest

from oracle import chunk_codecs as cc
from pinot_b200 import sql
from reduce_util import normalise

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name,num_docs,start", [("fixedByteRaw.v2", 2000, 100.2356), ("fixedByteCompressed.v2", 2000, 100.2356),
                                                 ("fixedByteSVRDoubles.v1", 10009, 0.0)])
def test_codecs_on_reference_written_files(name, num_docs, start):
    blob = np.fromfile(os.path.join(HERE, "golden", "raw_forward", name), dtype=np.uint8)
    got = np.frombuffer(cc.decode_fixed_byte_forward(blob, 8, num_docs), dtype=">f8")
    assert np.array_equal(got, np.arange(num_docs) + start)


def test_codec_round_trips():
    rng = np.random.default_rng(0)
    for trial in range(60):
        n = int(rng.integers(0, 3000))
        data = [rng.integers(0, 256, size=n, dtype=np.uint8), rng.integers(0, 3, size=n, dtype=np.uint8), np.zeros(n, dtype=np.uint8)][trial % 3].tobytes()
        assert cc.snappy_decode(cc.snappy_encode(data)) == data
        assert cc.lz4_block_decode(cc.lz4_block_encode(data)) == data


class _Values:
    def __init__(self, seg, q, r):
        self.seg, self.raw = seg, {q.group_by[j]: v for j, v in r.raw_key_values.items()}

    def value_of(self, c, i):
        return self.raw[c][i].item() if c in self.raw else self.seg.value_of(c, i)


def _table(seg, q, r):
    from gpu_util import oracle_table
    return oracle_table(seg, q, r)


def test_raw_group_by_equals_dictionary_group_by(oracle):
    rng = np.random.default_rng(5)
    n = 20_000
    cols = {"a": (rng.integers(0, 40, size=n) * 11 - 7).astype(np.int32), "b": (rng.integers(0, 30, size=n).astype(np.int64) * 5_000_000_000),
            "c": (rng.integers(0, 25, size=n) / 4.0).astype(np.float64), "f": (rng.integers(0, 9, size=n) / 2.0 - 1).astype(np.float32),
            "k": rng.integers(0, 6, size=n).astype(np.int32), "v": rng.integers(-50, 50, size=n).astype(np.int32)}
    raw = oracle.build_segment("r", cols, raw=["a", "b", "c", "f"], raw_compression={"a": cc.SNAPPY, "b": cc.LZ4, "c": cc.LZ4_LENGTH_PREFIXED})
    dic = oracle.build_segment("d", cols)
    for text in ("SELECT COUNT(*), SUM(v) FROM t GROUP BY a", "SELECT COUNT(*), MAX(v) FROM t WHERE v > 0 GROUP BY b",
                 "SELECT SUM(v), MIN(c) FROM t GROUP BY c, k", "SELECT COUNT(*) FROM t WHERE a > 100 GROUP BY k, a, f",
                 "SELECT SUM(a), AVG(c), MAX(b), MIN(f) FROM t WHERE b >= 50000000000 GROUP BY k",
                 "SELECT DISTINCTCOUNT(a), DISTINCTCOUNT(c), COUNT(*) FROM t WHERE k > 1",
                 "SELECT DISTINCTCOUNT(b), DISTINCTCOUNT(f) FROM t GROUP BY k", "SELECT DISTINCTCOUNT(c) FROM t WHERE a > 100 GROUP BY a"):
        q = sql.parse(text, num_groups_limit=1_000_000)
        r_raw, r_dic = oracle.execute(raw, q), oracle.execute(dic, q)
        if any(c in ("a", "b", "c", "f") for c in q.group_by):
            assert r_raw.regime == "NO_DICTIONARY"
        assert _table(raw, q, r_raw) == _table(dic, q, r_dic), text
        assert r_raw.stats[0] == r_dic.stats[0]


def test_raw_group_ids_are_first_seen_and_the_limit_cuts_in_doc_order(oracle):
    vals = np.asarray([50, 10, 50, 30, 10, 70, 90, 30, 110], dtype=np.int32)
    seg = oracle.build_segment("o", {"a": vals, "v": np.arange(len(vals), dtype=np.int32)}, raw=["a"])
    r = oracle.execute(seg, sql.parse("SELECT COUNT(*), SUM(v) FROM t GROUP BY a", num_groups_limit=1000))
    assert [r.raw_key_values[0][i] for i in r.keys[:, 0]] == [50, 10, 30, 70, 90, 110]       # first-seen order
    assert list(r.longs[0]) == [2, 2, 2, 1, 1, 1]
    r = oracle.execute(seg, sql.parse("SELECT COUNT(*), SUM(v) FROM t GROUP BY a", num_groups_limit=3))
    assert [r.raw_key_values[0][i] for i in r.keys[:, 0]] == [50, 10, 30] and r.groups_limit_reached   # 70, 90, 110 -> INVALID_ID
    assert list(r.longs[0]) == [2, 2, 2]
