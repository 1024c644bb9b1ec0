"""Format-level tests of the oracle, modelled on the reference's own unit tests:

* FixedBitIntReaderTest (pinot-segment-local/src/test/.../io/reader/impl/FixedBitIntReaderTest.java:51-83): every width
  1..31, random values, read / readUnchecked / read32 agree with what was written;
* FixedBitSVForwardIndexReaderV2Test (.../segment/index/readers/forward/FixedBitSVForwardIndexReaderV2Test.java:73-108):
  readDictIds over sequential, sparse and tail-of-buffer doc ids;
* golden BYTES of a reference-built v1 segment (paddingOld.tar.gz) for the forward index and the dictionaries;
* Roaring portable format round trips across array / bitmap / run containers (set equality -- byte parity is unpinned,
  the reference holds no golden bitmap bytes).
"""
import os

import numpy as np
import pytest

from oracle import segment_builder as sb

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden(name):
    return np.frombuffer(open(os.path.join(GOLDEN, name), "rb").read(), dtype=np.uint8)


@pytest.mark.parametrize("bits", range(1, 32))
def test_fixed_bit_reader_all_widths(oracle, bits):
    rng = np.random.default_rng(bits)
    n = 95 + 64  # FixedBitIntReaderTest uses 95 values; add two aligned groups for read32
    vals = rng.integers(0, 1 << bits, size=n, dtype=np.int64).astype(np.int32)
    buf = oracle.bitset_write(vals, bits)
    assert len(buf) == (n * bits + 7) // 8
    assert np.array_equal(buf, sb.pack_fixed_bits(vals, bits))  # numpy packer == byte-at-a-time writer
    padded = np.concatenate([buf, np.zeros(8, dtype=np.uint8)])
    for i in range(n):
        assert oracle.bitset_read(buf, i, bits) == vals[i]
        assert oracle.read_unchecked(padded, i, bits) == vals[i]
    for start in range(0, n - 31, 32):
        assert np.array_equal(oracle.read32(buf, start, bits), vals[start:start + 32])
    assert np.array_equal(sb.unpack_fixed_bits(buf, n, bits), vals)


@pytest.mark.parametrize("bits", [1, 3, 7, 8, 13, 16, 17, 20, 25, 26, 31])
def test_forward_index_read_dict_ids(oracle, bits):
    rng = np.random.default_rng(100 + bits)
    n = 99_999
    vals = rng.integers(0, 1 << bits, size=n, dtype=np.int64).astype(np.int32)
    buf = sb.pack_fixed_bits(vals, bits)
    # sequential from 32 different offsets, 10 000-doc blocks
    for off in range(0, 32, 5):
        ids = np.arange(off, min(n, off + 10_000), dtype=np.int32)
        assert np.array_equal(oracle.read_dict_ids(buf, n, bits, ids), vals[ids])
    # sparse stride 5..10
    ids = np.cumsum(rng.integers(5, 11, size=9_000)).astype(np.int32)
    ids = ids[ids < n]
    assert np.array_equal(oracle.read_dict_ids(buf, n, bits, ids), vals[ids])
    # tail of the buffer (the last two docs use the bounds-safe reader)
    ids = np.arange(n - 100, n, dtype=np.int32)
    assert np.array_equal(oracle.read_dict_ids(buf, n, bits, ids), vals[ids])
    ids = np.array([0, n - 2, n - 1], dtype=np.int32)
    assert np.array_equal(oracle.read_dict_ids(buf, n, bits, ids), vals[ids])


def test_num_bits_per_value(oracle):
    # PinotDataBitSet.getNumBitsPerValue javadoc examples (:49-59)
    for max_value, bits in [(0, 1), (1, 1), (2, 2), (9, 4), (113, 7), (255, 8), (256, 9), (9999, 14), (999_999, 20),
                            (2**31 - 1, 31)]:
        assert oracle.lib.po_num_bits_per_value(max_value) == bits
        assert sb.num_bits_per_value(max_value) == bits


def test_golden_bytes_padding_old_segment(oracle):
    """Real index files written by the reference (pinot-core/src/test/resources/data/paddingOld.tar.gz, 5 docs)."""
    # metadata.properties: age INT card 5 bits 3; outgoingName1 LONG card 5 bits 3 (start 246, end 902); percent FLOAT
    age_dict = _golden("padding_old_age_dict.bin")
    assert np.array_equal(np.frombuffer(age_dict.tobytes(), dtype=">i4"), [617, 824, 837, 1209, 1228])
    age_fwd = _golden("padding_old_age_sv_unsorted_fwd.bin")
    ids = [oracle.bitset_read(age_fwd, i, 3) for i in range(5)]
    assert sorted(ids) == [0, 1, 2, 3, 4]
    assert np.array_equal(oracle.bitset_write(np.array(ids, dtype=np.int32), 3), age_fwd)
    assert np.array_equal(sb.pack_fixed_bits(np.array(ids), 3), age_fwd)
    time_dict = np.frombuffer(_golden("padding_old_outgoingName1_dict.bin").tobytes(), dtype=">i8")
    assert time_dict[0] == 246 and time_dict[-1] == 902 and np.all(np.diff(time_dict) > 0)
    pct = np.frombuffer(_golden("padding_old_percent_dict.bin").tobytes(), dtype=">f4")
    assert np.all(np.diff(pct) > 0)  # sorted floats, big-endian
    # our builder reproduces the reference's bytes from the decoded values
    col = sb.build_column("age", np.array([617, 824, 837, 1209, 1228], dtype=np.int32)[ids])
    assert np.array_equal(col.dict, age_dict) and np.array_equal(col.fwd, age_fwd) and col.bits == 3
    name_dict = _golden("padding_old_name_dict.bin").tobytes()  # STRING, lengthOfEachEntry = 9, 2 entries
    assert len(name_dict) == 18


@pytest.mark.parametrize("run_optimize", [False, True])
def test_roaring_round_trip(oracle, run_optimize):
    rng = np.random.default_rng(7)
    cases = [
        np.array([], dtype=np.uint32),
        np.array([0], dtype=np.uint32),
        np.array([65535, 65536, 131071], dtype=np.uint32),
        np.arange(0, 5000, dtype=np.uint32),                                  # one run / bitmap container
        np.unique(rng.integers(0, 65536, size=3000)).astype(np.uint32),      # array container
        np.unique(rng.integers(0, 65536, size=40000)).astype(np.uint32),     # bitmap container
        np.unique(rng.integers(0, 10_000_000, size=200_000)).astype(np.uint32),
        np.concatenate([np.arange(100, 70_000), np.arange(200_000, 200_010), [4_000_000_000]]).astype(np.uint32),
        np.arange(0, 6 * 65536, 2, dtype=np.uint32),                          # >= 4 containers (offset header rule)
    ]
    for v in cases:
        buf = oracle.roaring_serialize(v, run_optimize)
        assert np.array_equal(oracle.roaring_deserialize(buf), v)
        if len(v):
            cookie = int(np.frombuffer(buf[:4].tobytes(), dtype="<u4")[0])
            assert cookie == 12346 or (cookie & 0xFFFF) == 12347


def test_inverted_index_layout(oracle):
    rng = np.random.default_rng(3)
    card, n = 37, 20_000
    ids = rng.integers(0, card, size=n).astype(np.int32)
    inv = oracle.inverted_index_build(ids, card)
    offs = np.frombuffer(inv[: 4 * (card + 1)].tobytes(), dtype=">u4")
    assert offs[0] == 4 * (card + 1) and offs[-1] == len(inv) and np.all(np.diff(offs.astype(np.int64)) > 0)
    for d in range(card):
        docs = oracle.roaring_deserialize(inv[offs[d]:offs[d + 1]])
        assert np.array_equal(docs, np.nonzero(ids == d)[0])


def test_segment_directory_writer_follows_the_reference_layouts(oracle, tmp_path):
    """tests/segment_dir_util.py (the helper the loader tests use) against what the reference wrote: the star-tree index
    map grammar of the reference-built fixture, the metadata key names, the v3 magic marker."""
    import os
    import re
    import struct

    import numpy as np

    from oracle import startree_builder as stb
    from segment_dir_util import write_segment_dir

    rng = np.random.default_rng(3)
    seg = oracle.build_segment("w", {"d1": rng.integers(0, 9, size=700).astype(np.int32),
                                     "m": rng.integers(0, 90, size=700).astype(np.int32)}, inverted=["d1"])
    st = stb.build_star_tree(seg, ["d1"], [("COUNT", None), ("MAX", "m")], max_leaf_records=10)
    root = write_segment_dir(str(tmp_path / "s"), seg, "v3", star_tree=st)
    here = os.path.dirname(os.path.abspath(__file__))
    grammar = re.compile(r"^(\d+)\.(.+)\.(STAR_TREE|FORWARD_INDEX)\.(OFFSET|SIZE) = (\d+)$")
    ref_lines = [l.strip() for l in open(os.path.join(here, "golden", "star_tree_index_map.txt")) if l.strip()]
    our_lines = [l.strip() for l in open(os.path.join(root, "v3", "star_tree_index_map")) if l.strip()]
    assert all(grammar.match(l) for l in ref_lines) and all(grammar.match(l) for l in our_lines)
    assert {grammar.match(l).group(2) for l in ref_lines} >= {"null", "count__*"}       # tree under "null", pair names
    assert {grammar.match(l).group(2) for l in our_lines} == {"null", "d1", "count__*", "max__m"}
    sizes = {grammar.match(l).group(2): int(grammar.match(l).group(5)) for l in our_lines if l.split(".")[-1].startswith("SIZE")}
    assert sum(sizes.values()) == os.path.getsize(os.path.join(root, "v3", "star_tree_index"))
    ref_meta = open(os.path.join(here, "golden", "star_tree_metadata.txt")).read()
    our_meta = open(os.path.join(root, "v3", "metadata.properties")).read()
    for key in ("startree.v2.count", "startree.v2.0.total.docs", "startree.v2.0.split.order", "startree.v2.0.function.column.pairs",
                "column.d1.cardinality", "column.d1.bitsPerElement", "column.d1.hasDictionary", "column.d1.isSorted"):
        assert key.replace("d1", "AirlineID") in ref_meta and key in our_meta, key
    psf = open(os.path.join(root, "v3", "columns.psf"), "rb").read()
    imap = dict(l.strip().split(" = ") for l in open(os.path.join(root, "v3", "index_map")) if " = " in l)
    for col in ("d1", "m"):
        off = int(imap[f"{col}.forward_index.startOffset"])
        assert struct.unpack(">Q", psf[off:off + 8])[0] == 0xDEADBEEFDEAFBEAD
    assert "d1.inverted_index.startOffset" in imap
