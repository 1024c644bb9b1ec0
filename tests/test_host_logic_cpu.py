"""CPU tests of the PRODUCT's host-side logic (inside libpinot_b200.so, no GPU needed) against the oracle:
star-tree traversal (pb200h::StarTree) and predicate -> dictId resolution (pb200h::matching_dict_ids).

The product functions are reached through a tiny test shim (tests/host_probe/host_probe.cpp) that is compiled with g++ and
linked against the in-tree library; the library loads without a GPU (only pb200_init needs one)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import segment_builder as sb
from oracle import startree_builder as stb
from oracle.startree_query import matching_dict_ids as oracle_matching_dict_ids
from pinot_b200 import _lib
from pinot_b200.plan_maker import _marshal_query
from pinot_b200.query import Aggregation, Predicate, QueryContext

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def probe():
    from pinot_b200.build import build
    build()
    src = os.path.join(HERE, "host_probe", "host_probe.cpp")
    out = os.path.join(HERE, "host_probe", "libhost_probe.so")
    libdir = os.path.join(ROOT, "pinot_b200")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(_lib.LIB_PATH)):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", f"-I{cuda_inc}", "-o", out, src, f"-L{libdir}",
                        "-lpinot_b200", f"-Wl,-rpath,{libdir}"], check=True)
    lib = C.CDLL(out)
    lib.probe_startree_traverse.restype = C.c_int64
    lib.probe_startree_traverse.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_uint32, C.c_void_p, C.c_int64, C.POINTER(C.c_uint32)]
    lib.probe_matching_dict_ids.restype = C.c_int64
    lib.probe_matching_dict_ids.argtypes = [C.POINTER(_lib.HColumn), C.POINTER(_lib.HFilterNode), C.POINTER(_lib.HLiteral),
                                            C.c_void_p, C.c_int64]
    lib.probe_decode_fixed_byte_forward.restype = C.c_int32
    lib.probe_decode_fixed_byte_forward.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_int64, C.c_void_p]
    lib.probe_synthesize_dictionary.restype = C.c_int32
    lib.probe_synthesize_dictionary.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                                                C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    return lib


def _product_traverse(probe, tree, num_dims, preds, gb_dims, max_docs):
    dims = np.asarray(sorted(preds), dtype=np.int32)
    arrays = [np.sort(np.asarray(preds[d], dtype=np.int32)) for d in dims]
    offsets = np.concatenate([[0], np.cumsum([len(a) for a in arrays])]).astype(np.int32)
    ids = np.concatenate(arrays).astype(np.int32) if arrays else np.zeros(0, dtype=np.int32)
    ids = np.ascontiguousarray(np.concatenate([ids, [0]]).astype(np.int32))  # never a NULL pointer
    mask = 0
    for d in gb_dims:
        mask |= 1 << d
    out = np.zeros(2 * (max_docs + 1), dtype=np.int32)
    rem = C.c_uint32(0)
    n = probe.probe_startree_traverse(tree.ctypes.data, len(tree), num_dims, len(dims), dims.ctypes.data if len(dims) else None,
                                      offsets.ctypes.data, ids.ctypes.data, mask, out.ctypes.data, max_docs + 1, C.byref(rem))
    assert n != -2, "product rejected the tree"
    if n == -1:
        return None, []
    docs = set()
    for s, e in out[: 2 * n].reshape(-1, 2):
        docs.update(range(int(s), int(e)))
    return docs, [d for d in range(32) if rem.value >> d & 1]


def _compare_traversals(oracle, probe, tree, num_dims, num_docs, cases):
    for preds, gb in cases:
        want_docs, want_rem = oracle.startree_traverse(tree, preds, gb, num_docs)
        got_docs, got_rem = _product_traverse(probe, tree, num_dims, preds, gb, num_docs)
        if want_docs is None:
            assert got_docs is None, (preds, gb)
            continue
        assert got_docs == set(int(d) for d in want_docs), (preds, gb)
        assert sorted(got_rem) == sorted(want_rem), (preds, gb)


def test_product_traversal_on_the_reference_built_star_tree(oracle, probe):
    blob = np.fromfile(os.path.join(HERE, "golden", "star_tree_index.bin"), dtype=np.uint8)
    tree = np.ascontiguousarray(blob[:18715])  # star_tree_index_map: 0.null.STAR_TREE.OFFSET = 0, SIZE = 18715
    dims, _ = oracle.startree_info(tree)
    assert dims == ["AirlineID", "Origin", "Dest"]
    cases = [({}, []), ({}, [0]), ({}, [1]), ({}, [0, 2]), ({0: [3]}, []), ({0: [3, 5, 9]}, [1]), ({1: [10, 11, 12]}, []),
             ({0: [1], 2: [50, 51]}, []), ({2: list(range(0, 104, 2))}, [0]), ({0: list(range(14)), 1: [7]}, [2]),
             ({0: []}, []), ({1: [96], 2: [103]}, [0, 1, 2])]
    _compare_traversals(oracle, probe, tree, 3, 1004, cases)


@pytest.mark.parametrize("max_leaf", [1, 7, 100, 100000])
def test_product_traversal_on_built_trees(oracle, probe, max_leaf):
    rng = np.random.default_rng(max_leaf)
    n = 6000
    seg = oracle.build_segment("t", {"d1": rng.integers(0, 40, size=n).astype(np.int32),
                                     "d2": rng.integers(0, 25, size=n).astype(np.int32),
                                     "d3": rng.integers(0, 4, size=n).astype(np.int32),
                                     "m": rng.integers(0, 99, size=n).astype(np.int32)})
    st = stb.build_star_tree(seg, ["d1", "d2", "d3"], [("COUNT", None), ("SUM", "m")], max_leaf_records=max_leaf)
    cases = []
    for _ in range(40):
        preds = {}
        for d, card in ((0, 40), (1, 25), (2, 4)):
            r = rng.random()
            if r < 0.35:
                preds[d] = rng.choice(card, size=int(rng.integers(1, max(2, card // 2))), replace=False).tolist()
            elif r < 0.42:
                preds[d] = list(range(card))           # matches every value: the star node may stand in
            elif r < 0.45:
                preds[d] = []                          # matches nothing
        gb = [d for d in range(3) if rng.random() < 0.3]
        cases.append((preds, gb))
    _compare_traversals(oracle, probe, st.tree, 3, st.num_docs, cases)


def _column_struct(col):
    nm = col.name.encode()
    return _lib.HColumn(nm, col.data_type, 1, col.bits, col.cardinality, int(col.is_sorted), col.dict_entry_bytes,
                        col.fwd.ctypes.data, len(col.fwd), col.dict.ctypes.data, len(col.dict), None, 0), nm


def test_product_predicate_resolution_matches_the_oracle(oracle, probe):
    rng = np.random.default_rng(5)
    n = 3000
    seg = oracle.build_segment("p", {
        "i": rng.integers(-500, 500, size=n).astype(np.int32) * 3,
        "l": rng.integers(0, 400, size=n).astype(np.int64) * 10_000_000_019,
        "f": (rng.integers(0, 300, size=n) / 8.0).astype(np.float32),
        "d": (rng.integers(-200, 200, size=n) / 3.0).astype(np.float64),
        "s": np.array([b"aa", b"ab", b"b", b"ba", b"zz", b"m", b"mm"])[rng.integers(0, 7, size=n)],
    })
    checked = 0
    for col in seg.columns:
        vals = col.dict_values
        present = [v.decode() if isinstance(v, bytes) else v.item() for v in vals]
        if col.data_type == sb.STRING:
            absent = ["", "a", "abc", "zzz", "n"]
        elif col.data_type in (sb.INT, sb.LONG):
            absent = [int(present[0]) - 1, int(present[-1]) + 1, int(present[3]) + 1]
        else:
            absent = [float(present[0]) - 0.5, float(present[-1]) + 0.5, float(present[3]) + 1e-3]
        pool = present[:: max(1, len(present) // 12)] + absent
        preds = []
        for v in pool:
            preds += [Predicate("EQ", col.name, [v]), Predicate("NEQ", col.name, [v])]
        for _ in range(12):
            k = int(rng.integers(1, 6))
            vs = [pool[int(j)] for j in rng.integers(0, len(pool), size=k)]
            preds += [Predicate("IN", col.name, vs), Predicate("NOT_IN", col.name, vs)]
        for lo in pool[:8] + [None]:
            for hi in pool[-8:] + [None]:
                for li in (True, False):
                    for ui in (True, False):
                        preds.append(Predicate("RANGE", col.name, [], lo, hi, li, ui))
        cstruct, _keep = _column_struct(col)
        for p in preds:
            if p.type == "RANGE" and p.lower is not None and p.upper is not None and p.lower > p.upper:
                continue
            hq, keep = _marshal_query(QueryContext([Aggregation("COUNT", None)], filter=p), False)
            out = np.zeros(col.cardinality + 1, dtype=np.int32)
            m = probe.probe_matching_dict_ids(C.byref(cstruct), hq.filter, hq.literals, out.ctypes.data, len(out))
            want = oracle_matching_dict_ids(col, p)
            assert m == len(want) and np.array_equal(out[:m], want), (col.name, p)
            checked += 1
    assert checked > 1500


F_MATCH_ALL, F_EMPTY, F_SCAN_RANGE, F_SCAN_IN, F_SCAN_NOT_IN, F_INV_IN, F_INV_NOT_IN, F_DOC_RANGES = 3, 4, 5, 6, 7, 8, 9, 10


def _leaf_docs(op, lo, hi, ids, dict_ids, num_docs):
    """What the device leaf the product built would match, evaluated with numpy on the column's dictIds."""
    if op == F_MATCH_ALL:
        return np.arange(num_docs)
    if op == F_EMPTY:
        return np.zeros(0, dtype=np.int64)
    if op == F_SCAN_RANGE:
        return np.nonzero((dict_ids >= lo) & (dict_ids < hi))[0]
    if op in (F_SCAN_IN, F_INV_IN):
        return np.nonzero(np.isin(dict_ids, ids))[0]
    if op in (F_SCAN_NOT_IN, F_INV_NOT_IN):
        return np.nonzero(~np.isin(dict_ids, ids))[0]
    assert op == F_DOC_RANGES, op
    m = np.zeros(num_docs, dtype=bool)
    for s, e in np.asarray(ids).reshape(-1, 2):
        m[s:e + 1] = True  # inclusive pairs
    return np.nonzero(m)[0]


def test_product_leaf_choice_and_matches(oracle, probe):
    """pb200h::leaf_to_device = PredicateEvaluator resolution + FilterOperatorUtils.getLeafFilterOperator: the leaf kind follows
    the reference's priorities (sorted index, then inverted index for EQ/IN-like predicates, else scan) and the docs it would
    match are exactly the oracle's docs for that predicate."""
    probe.probe_leaf_to_device.restype = C.c_int64
    probe.probe_leaf_to_device.argtypes = [C.POINTER(_lib.HColumn), C.c_int32, C.POINTER(_lib.HFilterNode),
                                           C.POINTER(_lib.HLiteral), C.c_void_p, C.c_void_p, C.c_int64]
    rng = np.random.default_rng(9)
    n = 5000
    seg = oracle.build_segment("lf", {
        "plain": rng.integers(0, 200, size=n).astype(np.int32) * 5,
        "inv": rng.integers(0, 60, size=n).astype(np.int32),
        "srt": np.sort(rng.integers(0, 30, size=n)).astype(np.int32) * 2,
        "str": np.array([b"k1", b"k2", b"k3", b"k9"])[rng.integers(0, 4, size=n)],
    }, inverted=["inv", "str"])
    checked = 0
    for col in seg.columns:
        present = [v.decode() if isinstance(v, bytes) else v.item() for v in col.dict_values]
        absent = ["k0", "k5", "zz"] if col.data_type == sb.STRING else [present[0] - 1, present[-1] + 7, present[2] + 1]
        pool = present[:: max(1, len(present) // 10)] + absent
        preds = [Predicate(t, col.name, [v]) for v in pool for t in ("EQ", "NEQ")]
        for _ in range(10):
            vs = [pool[int(j)] for j in rng.integers(0, len(pool), size=int(rng.integers(1, 5)))]
            preds += [Predicate("IN", col.name, vs), Predicate("NOT_IN", col.name, vs)]
        for lo in pool[:5] + [None]:
            for hi in pool[-5:] + [None]:
                for li, ui in ((True, True), (False, True), (True, False), (False, False)):
                    if lo is None or hi is None or lo <= hi:
                        preds.append(Predicate("RANGE", col.name, [], lo, hi, li, ui))
        nm = col.name.encode()
        has_inv = col.inv is not None and not col.is_sorted
        cstruct = _lib.HColumn(nm, col.data_type, 1, col.bits, col.cardinality, int(col.is_sorted), col.dict_entry_bytes,
                               col.fwd.ctypes.data, len(col.fwd), col.dict.ctypes.data, len(col.dict),
                               col.inv.ctypes.data if has_inv else None, len(col.inv) if has_inv else 0)
        for p in preds:
            q = QueryContext([Aggregation("COUNT", None)], filter=p)
            hq, keep = _marshal_query(q, False)
            out4 = np.zeros(4, dtype=np.int32)
            ids = np.zeros(2 * col.cardinality + 8, dtype=np.int32)
            m = probe.probe_leaf_to_device(C.byref(cstruct), seg.num_docs, hq.filter, hq.literals, out4.ctypes.data,
                                           ids.ctypes.data, len(ids))
            assert m >= 0, (col.name, p, m)
            op, lo, hi = int(out4[0]), int(out4[2]), int(out4[3])
            if op not in (F_MATCH_ALL, F_EMPTY):  # FilterOperatorUtils.getLeafFilterOperator :74-133
                if col.is_sorted:
                    assert op == F_DOC_RANGES, (col.name, p, op)
                elif p.type == "RANGE":
                    assert op == F_SCAN_RANGE, (col.name, p, op)
                elif has_inv:
                    assert op in (F_INV_IN, F_INV_NOT_IN), (col.name, p, op)
                else:
                    assert op in (F_SCAN_IN, F_SCAN_NOT_IN), (col.name, p, op)
            got = _leaf_docs(op, lo, hi, ids[:m], col.dict_ids, seg.num_docs)
            want, _ = oracle.filter_doc_ids(seg, q)
            assert np.array_equal(got, want), (col.name, p, op)
            checked += 1
    assert checked > 600


def test_product_plan_constant_folding_and_operator_choice(oracle, probe):
    """pb200h_explain on a host-only segment: predicates that resolve to always-true / always-false are folded through
    AND / OR / NOT exactly like FilterPlanNode + FilterOperatorUtils do, and the operator choice follows AggregationPlanNode
    (empty filter -> EMPTY, match-all + dictionary-answerable functions -> NonScanBasedAggregationOperator)."""
    from pinot_b200 import sql
    probe.probe_explain.restype = C.c_int32
    probe.probe_explain.argtypes = [C.POINTER(_lib.HColumn), C.c_int32, C.c_int32, C.POINTER(_lib.HQuery), C.c_char_p, C.c_int32]
    rng = np.random.default_rng(2)
    n = 2000
    seg = oracle.build_segment("pl", {"a": rng.integers(0, 50, size=n).astype(np.int32),
                                      "b": rng.integers(0, 10, size=n).astype(np.int32),
                                      "s": np.sort(rng.integers(0, 8, size=n)).astype(np.int32)}, inverted=["b"])
    arr = (_lib.HColumn * len(seg.columns))()
    keep = []
    for i, c in enumerate(seg.columns):
        nm = c.name.encode()
        keep.append(nm)
        has_inv = c.inv is not None and not c.is_sorted
        arr[i] = _lib.HColumn(nm, c.data_type, 1, c.bits, c.cardinality, int(c.is_sorted), c.dict_entry_bytes, c.fwd.ctypes.data,
                              len(c.fwd), c.dict.ctypes.data, len(c.dict), c.inv.ctypes.data if has_inv else None,
                              len(c.inv) if has_inv else 0)

    def explain(text):
        hq, k = _marshal_query(sql.parse(text), False)
        buf = C.create_string_buffer(4096)
        m = probe.probe_explain(arr, len(seg.columns), seg.num_docs, C.byref(hq), buf, 4096)
        assert m >= 0, text
        return buf.value.decode()

    assert explain("SELECT SUM(a) FROM t WHERE a > 1000") == "EMPTY(FILTER_EMPTY)"                       # beyond the dictionary
    assert explain("SELECT SUM(a) FROM t WHERE a > 1000 AND b = 3") == "EMPTY(FILTER_EMPTY)"             # AND with always-false
    assert explain("SELECT SUM(a) FROM t WHERE a >= 0") == "AGGREGATE(FILTER_MATCH_ENTIRE_SEGMENT)"      # always-true
    assert explain("SELECT MAX(a), MIN(b) FROM t WHERE a >= 0 OR b = 3") == "AGGREGATE_NO_SCAN(FILTER_MATCH_ENTIRE_SEGMENT)"
    assert explain("SELECT COUNT(*) FROM t") == "AGGREGATE_NO_SCAN(FILTER_MATCH_ENTIRE_SEGMENT)"
    # partial constants stay in the tree: every segment of a submission must keep the same tree SHAPE (one launch for all
    # segments), so an always-false operand becomes an EMPTY leaf instead of being dropped as FilterOperatorUtils would
    assert explain("SELECT SUM(a) FROM t WHERE a > 1000 OR b = 3") == "AGGREGATE(FILTER_OR(FILTER_INVERTED_INDEX(EQ,b),FILTER_EMPTY))"
    assert explain("SELECT SUM(a) FROM t WHERE NOT (a > 1000)") == "AGGREGATE(FILTER_MATCH_ENTIRE_SEGMENT)"
    g = explain("SELECT SUM(a) FROM t WHERE b != 3 AND s = 2 AND a < 10 GROUP BY b")
    assert g.startswith("GROUP_BY(FILTER_AND(") and "FILTER_SORTED_INDEX" in g and "FILTER_INVERTED_INDEX" in g and "FILTER_FULL_SCAN(RANGE,a,[0,10))" in g


# ---- raw (no-dictionary) forward indexes at load: pinot_b200/csrc/host/raw_forward.cpp ------------------------------------
RAW_FIXTURES = [("fixedByteRaw.v2", 2000, 100.2356), ("fixedByteCompressed.v2", 2000, 100.2356), ("fixedByteSVRDoubles.v1", 10009, 0.0)]


def _product_decode(probe, file_bytes, width, n):
    buf = np.ascontiguousarray(np.frombuffer(bytes(file_bytes), dtype=np.uint8))
    out = np.zeros(n * width, dtype=np.uint8)
    rc = probe.probe_decode_fixed_byte_forward(buf.ctypes.data, len(buf), width, n, out.ctypes.data)
    return rc, out


@pytest.mark.parametrize("name,num_docs,start", RAW_FIXTURES)
def test_product_decodes_the_reference_written_raw_forward_indexes(probe, name, num_docs, start):
    """FixedByteChunkSVForwardIndexTest.java:340-377 (testBackwardCompatibilityV1 / V2): value i == i + startValue, read from
    files old versions of the reference wrote (PASS_THROUGH v2, SNAPPY v2, SNAPPY v1)."""
    from oracle import chunk_codecs as cc
    blob = np.fromfile(os.path.join(HERE, "golden", "raw_forward", name), dtype=np.uint8)
    rc, out = _product_decode(probe, blob, 8, num_docs)
    assert rc == 0
    want = np.arange(num_docs) + start
    assert np.array_equal(out.view(">f8"), want)
    assert np.array_equal(np.frombuffer(cc.decode_fixed_byte_forward(blob, 8, num_docs), dtype=">f8"), want)  # the oracle's codec too


@pytest.mark.parametrize("compression", [0, 1, 3, 4])
@pytest.mark.parametrize("version", [2, 3, 4])
def test_product_chunk_decoders_equal_the_oracle_codecs(probe, compression, version):
    """PASS_THROUGH / SNAPPY / LZ4 / LZ4_LENGTH_PREFIXED chunks, int and long chunk offsets, ragged last chunk."""
    from oracle import chunk_codecs as cc
    rng = np.random.default_rng(100 * version + compression)
    for dt, n in ((">i4", 2501), (">i8", 1000), (">f4", 999), (">f8", 4097), (">i4", 1)):
        kind = rng.integers(0, 3)
        vals = (rng.integers(0, 40, size=n) if kind == 0 else rng.integers(-2**31, 2**31 - 1, size=n) if kind == 1
                else np.repeat(rng.integers(0, 1000, size=n // 7 + 1), 7)[:n])
        body = vals.astype(dt).tobytes()
        width = np.dtype(dt).itemsize
        f = cc.encode_fixed_byte_forward(body, width, n, compression, version, docs_per_chunk=int(rng.choice([100, 1000, 4096])))
        rc, out = _product_decode(probe, f, width, n)
        assert rc == 0 and out.tobytes() == body == cc.decode_fixed_byte_forward(f, width, n), (dt, n, compression, version)
    # truncated file: an error, not a crash or silent garbage
    body = np.arange(3000).astype(">f8").tobytes()
    f = cc.encode_fixed_byte_forward(body, 8, 3000, compression, version, docs_per_chunk=1000)
    assert _product_decode(probe, f[: len(f) - 9], 8, 3000)[0] != 0
    width, n = 8, 3000
    # ZSTANDARD / GZIP: refused
    z = f.copy(); z[20:24] = [0, 0, 0, 2]
    assert _product_decode(probe, z, width, n)[0] != 0


@pytest.mark.parametrize("np_type,data_type", [(np.int32, sb.INT), (np.int64, sb.LONG), (np.float32, sb.FLOAT), (np.float64, sb.DOUBLE)])
def test_product_dictionary_synthesis_equals_a_sorted_unique_dictionary(oracle, probe, np_type, data_type):
    """A raw column with few distinct values becomes dictionary + fixed-bit forward index at load: the dictionary must be the
    sorted distinct values (SegmentDictionaryCreator order) and the forward index the FixedBitSVForwardIndexWriter layout."""
    rng = np.random.default_rng(int(data_type) + 5)
    for card, n in ((1, 10), (2, 33), (3, 1000), (257, 5000), (4000, 4000)):
        pool = (rng.integers(-10**9, 10**9, size=card) * (1 if np_type in (np.int32,) else 3)).astype(np_type)
        if np_type in (np.float32, np.float64):
            pool = (pool / 7.0).astype(np_type)
            pool[0] = -0.0 if card > 1 else pool[0]
        vals = pool[rng.integers(0, card, size=n)]
        be = {np.int32: ">i4", np.int64: ">i8", np.float32: ">f4", np.float64: ">f8"}[np_type]
        body = np.ascontiguousarray(np.frombuffer(vals.astype(be).tobytes(), dtype=np.uint8))
        width = np.dtype(be).itemsize
        dict_out = np.zeros(n * width, dtype=np.uint8)
        fwd_out = np.zeros(n * 4 + 8, dtype=np.uint8)
        c, b = C.c_int32(), C.c_int32()
        assert probe.probe_synthesize_dictionary(body.ctypes.data, data_type, n, 1 << 20, dict_out.ctypes.data, len(dict_out),
                                                 fwd_out.ctypes.data, len(fwd_out), C.byref(c), C.byref(b)) == 1
        uniq, inv = np.unique(vals, return_inverse=True)
        assert c.value == len(uniq) and b.value == sb.num_bits_per_value(len(uniq) - 1)
        assert np.array_equal(dict_out[: c.value * width].view(be).astype(np_type), uniq)
        ids = oracle.read_dict_ids(fwd_out, n, b.value, np.arange(n, dtype=np.int32))
        assert np.array_equal(ids, inv.astype(np.int32))
        # too many distinct values: stays raw
        if len(uniq) > 2:
            assert probe.probe_synthesize_dictionary(body.ctypes.data, data_type, n, len(uniq) - 1, dict_out.ctypes.data, len(dict_out),
                                                     fwd_out.ctypes.data, len(fwd_out), C.byref(c), C.byref(b)) == 0


def test_product_chunk_decoders_survive_malformed_input(probe):
    """Memory safety of the load-time decoders: random bytes, truncated and bit-flipped valid files must end in an error or
    in decoded bytes -- never in a crash (the test process would die) or an out-of-bounds write (guard bytes stay intact)."""
    from oracle import chunk_codecs as cc
    rng = np.random.default_rng(77)
    body = np.repeat(rng.integers(0, 50, size=600), 5)[:2500].astype(">i8").tobytes()
    valid = [cc.encode_fixed_byte_forward(body, 8, 2500, comp, ver, docs_per_chunk=700) for comp in (0, 1, 3, 4) for ver in (2, 3)]
    valid.append(cc.encode_fixed_byte_forward(body, 8, 2500, cc.SNAPPY, 1, docs_per_chunk=700))
    cases = [rng.integers(0, 256, size=int(rng.integers(0, 400)), dtype=np.uint8) for _ in range(200)]
    for f in valid:
        for _ in range(60):
            g = f.copy()
            k = int(rng.integers(1, 6))
            for pos in rng.integers(0, len(g), size=k):
                g[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))
            cases.append(g)
        for cut in rng.integers(0, len(f), size=20):
            cases.append(f[: int(cut)].copy())
    ok = bad = 0
    for g in cases:
        buf = np.ascontiguousarray(np.concatenate([g, np.zeros(1, dtype=np.uint8)]))   # never a NULL pointer
        out = np.full(2500 * 8 + 64, 0xA5, dtype=np.uint8)
        rc = probe.probe_decode_fixed_byte_forward(buf.ctypes.data, len(g), 8, 2500, out.ctypes.data)
        assert np.all(out[2500 * 8:] == 0xA5)   # nothing written past the value array
        ok += rc == 0
        bad += rc != 0
    assert bad > 100 and ok > 0   # both outcomes occur (bit flips inside literals still decode)


def test_star_tree_reader_survives_malformed_input(oracle, probe):
    """Truncated / bit-flipped OffHeapStarTree buffers: rejected (-2) or traversed, never a crash or an endless walk."""
    blob = np.fromfile(os.path.join(HERE, "golden", "star_tree_index.bin"), dtype=np.uint8)
    tree = np.ascontiguousarray(blob[:18715])
    rng = np.random.default_rng(5)
    cases = [tree[: int(c)].copy() for c in rng.integers(0, len(tree), size=40)]
    for _ in range(300):
        g = tree.copy()
        for pos in rng.integers(0, len(g), size=int(rng.integers(1, 4))):
            g[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))
        cases.append(g)
    dims = np.asarray([0], dtype=np.int32)
    offsets = np.asarray([0, 2], dtype=np.int32)
    ids = np.asarray([1, 3, 0], dtype=np.int32)
    out = np.zeros(2 * 200_001, dtype=np.int32)
    rejected = walked = 0
    for g in cases:
        buf = np.ascontiguousarray(np.concatenate([g, np.zeros(1, dtype=np.uint8)]))
        rem = C.c_uint32(0)
        n = probe.probe_startree_traverse(buf.ctypes.data, len(g), 3, 1, dims.ctypes.data, offsets.ctypes.data, ids.ctypes.data, 0b110,
                                          out.ctypes.data, 200_000, C.byref(rem))
        rejected += n == -2
        walked += n >= -1
    assert rejected > 40 and walked > 0


def test_roaring_validation_before_the_device_reads_posting_lists(oracle):
    """pb200_roaring_validate (applied to every posting list at pb200_segment_register): bitmaps written by the oracle's
    serializer (array, bitmap and run containers, with and without the offset header) pass; a doc id beyond numDocs, truncated
    and bit-flipped buffers are rejected or accepted, never crash the process."""
    lib = _lib.load()
    rng = np.random.default_rng(31)
    num_docs = 300_000
    shapes = [np.sort(rng.choice(num_docs, size=k, replace=False)).astype(np.uint32) for k in (1, 7, 4096, 5000, 70_000, 200_000)]
    shapes.append(np.arange(1000, 250_000, dtype=np.uint32))                                  # runs
    shapes.append(np.concatenate([np.arange(10, 60_000), np.arange(70_000, 70_010), np.arange(140_000, 200_000)]).astype(np.uint32))
    shapes.append(np.zeros(0, dtype=np.uint32))
    good = []
    for vals in shapes:
        for run_opt in (0, 1):
            blob = oracle.roaring_serialize(vals, bool(run_opt))
            good.append((np.ascontiguousarray(blob), vals))
    for blob, vals in good:
        assert lib.pb200_roaring_validate(blob.ctypes.data, len(blob), num_docs) == 0
        if len(vals):
            assert lib.pb200_roaring_validate(blob.ctypes.data, len(blob), int(vals[-1])) != 0    # largest doc id == numDocs: out of range
            assert lib.pb200_roaring_validate(blob.ctypes.data, len(blob), int(vals[-1]) + 1) == 0
    rejected = 0
    for blob, vals in good:
        if len(blob) < 9:
            continue
        for _ in range(80):
            g = blob.copy()
            if rng.integers(0, 3) == 0:
                g = g[: int(rng.integers(1, len(g)))].copy()
            else:
                for pos in rng.integers(0, len(g), size=int(rng.integers(1, 4))):
                    g[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))
            buf = np.ascontiguousarray(np.concatenate([g, np.zeros(1, dtype=np.uint8)]))
            rc = lib.pb200_roaring_validate(buf.ctypes.data, len(g), num_docs)
            rejected += rc != 0
    assert rejected > 200
