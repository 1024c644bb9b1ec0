"""Star-tree builder -- TEST INFRASTRUCTURE (part of the oracle).

Restates ``BaseSingleTreeBuilder.build`` (``pinot-segment-local/.../startree/v2/builder/BaseSingleTreeBuilder.java:303-460``)
and ``StarTreeBuilderUtils.serializeTree`` (``.../startree/StarTreeBuilderUtils.java:130-230``):

1. sort the segment's records by the dimension split order (dictIds) and aggregate duplicates -> star-tree docs [0, n0);
2. ``constructStarTree``: per node split the doc range on the next dimension; when a node has more than one child and
   the dimension is not in ``skipStarNodeCreationForDimensions`` add a STAR child whose docs are the range aggregated over
   the remaining dimensions (the starred dimension stored as ``STAR_IN_FORWARD_INDEX`` = 0); recurse while a child holds
   more than ``maxLeafRecords`` docs;
3. ``createAggregatedDocs``: every node gets an aggregated doc (single-doc leaf: itself; star child present: the star
   child's; otherwise the merge of the children's, appended);
4. serialize breadth first, children sorted by dimension value (STAR = -1 first), LITTLE-endian, 7 ints per node;
   dimension forward indexes fixed-bit (bits of the base column), metric forward indexes raw PASS_THROUGH chunks
   (COUNT -> LONG, SUM / MIN / MAX -> DOUBLE: ``ValueAggregatorFactory.getAggregatedValueType``).

Doc-id assignment follows the same append order as the reference except where Java iterates a ``HashMap`` (children
order); the reader / traversal (the path under test) does not depend on it.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import segment_builder as sb

MAGIC = 0xBADDA55B00DAD00D
STAR = -1
STAR_IN_FORWARD_INDEX = 0


@dataclass
class StarTreeData:
    dimensions: List[str]
    function_pairs: List[Tuple[str, Optional[str]]]  # ("COUNT", None), ("SUM", "m"), ("MAX", "m") ...
    num_docs: int
    tree: np.ndarray  # serialized OffHeapStarTree bytes
    dim_dict_ids: np.ndarray  # [num_docs, D] int32
    metrics: List[np.ndarray]  # per pair: int64 (COUNT) or float64
    segment: sb.SegmentData = None  # star-tree docs as a segment: dims (dict-encoded) + metric columns (raw)
    max_leaf_records: int = 10000

    def metric_name(self, i: int) -> str:
        fn, col = self.function_pairs[i]
        return f"{fn.lower()}__{col if col else '*'}"


class _Node:
    __slots__ = ("dim", "value", "start", "end", "agg", "children")

    def __init__(self, dim=-1, value=-1, start=-1, end=-1):
        self.dim, self.value, self.start, self.end, self.agg, self.children = dim, value, start, end, -1, None


class _Docs:
    def __init__(self, d: int, kinds: Sequence[str]):
        self.dims = np.zeros((1024, d), dtype=np.int32)
        self.mets = [np.zeros(1024, dtype=np.int64 if k == "COUNT" else np.float64) for k in kinds]
        self.n = 0

    def append(self, dims: np.ndarray, mets: Sequence[np.ndarray]) -> int:
        k = dims.shape[0]
        while self.n + k > self.dims.shape[0]:
            self.dims = np.concatenate([self.dims, np.zeros_like(self.dims)])
            self.mets = [np.concatenate([m, np.zeros_like(m)]) for m in self.mets]
        self.dims[self.n:self.n + k] = dims
        for m, v in zip(self.mets, mets):
            m[self.n:self.n + k] = v
        start = self.n
        self.n += k
        return start


def _aggregate_sorted(dims: np.ndarray, mets: Sequence[np.ndarray], kinds: Sequence[str], key_cols: Sequence[int]):
    """dims sorted by key_cols: merge rows with identical key columns (mergeStarTreeRecord)."""
    n = dims.shape[0]
    if n == 0:
        return dims, list(mets)
    if key_cols:
        change = np.any(dims[1:, key_cols] != dims[:-1, key_cols], axis=1)
        starts = np.concatenate([[0], np.nonzero(change)[0] + 1])
    else:
        starts = np.array([0])
    out_d = dims[starts]
    out_m = []
    for m, k in zip(mets, kinds):
        if k in ("COUNT", "SUM"):
            out_m.append(np.add.reduceat(m, starts))
        elif k == "MAX":
            out_m.append(np.maximum.reduceat(m, starts))
        else:
            out_m.append(np.minimum.reduceat(m, starts))
    return out_d, out_m


def build_star_tree(seg: sb.SegmentData, dimensions: Sequence[str], function_pairs: Sequence[Tuple[str, Optional[str]]],
                    max_leaf_records: int = 10000, skip_star_for: Sequence[str] = ()) -> StarTreeData:
    D = len(dimensions)
    kinds = [fn for fn, _ in function_pairs]
    n = seg.num_docs
    dims = np.stack([seg.column(d).dict_ids for d in dimensions], axis=1).astype(np.int32)
    mets = []
    for fn, col in function_pairs:
        if fn == "COUNT":
            mets.append(np.ones(n, dtype=np.int64))
        else:
            c = seg.column(col)
            vals = c.dict_values[c.dict_ids] if c.has_dictionary else c.raw_values
            mets.append(vals.astype(np.float64))
    # 1. sortAndAggregateSegmentRecords
    order = np.lexsort([dims[:, j] for j in reversed(range(D))])
    dims, mets = dims[order], [m[order] for m in mets]
    dims, mets = _aggregate_sorted(dims, mets, kinds, list(range(D)))
    docs = _Docs(D, kinds)
    docs.append(dims, mets)
    n0 = docs.n
    skip = {dimensions.index(d) for d in skip_star_for}
    root = _Node()
    num_nodes = [1]

    def construct(node: _Node, start: int, end: int):
        child_dim = node.dim + 1
        if child_dim == D:
            return
        vals = docs.dims[start:end, child_dim]
        bounds = np.concatenate([[0], np.nonzero(vals[1:] != vals[:-1])[0] + 1, [end - start]])
        children = [_Node(child_dim, int(vals[bounds[i]]), start + int(bounds[i]), start + int(bounds[i + 1]))
                    for i in range(len(bounds) - 1)]
        num_nodes[0] += len(children)
        if child_dim not in skip and len(children) > 1:  # constructStarNode
            d = docs.dims[start:end].copy()
            m = [x[start:end].copy() for x in docs.mets]
            d[:, child_dim] = STAR_IN_FORWARD_INDEX
            rest = list(range(child_dim + 1, D))
            if rest:
                o = np.lexsort([d[:, j] for j in reversed(rest)])
                d, m = d[o], [x[o] for x in m]
            d, m = _aggregate_sorted(d, m, kinds, rest)
            s = docs.append(d, m)
            star = _Node(child_dim, STAR, s, docs.n)
            num_nodes[0] += 1
            children = [star] + children
        node.children = children
        for c in children:
            if c.end - c.start > max_leaf_records:
                construct(c, c.start, c.end)

    construct(root, 0, n0)

    def aggregated(node: _Node):
        """Returns (dims row, metric values) of the node's aggregated doc and sets node.agg."""
        if node.children is None:
            if node.start == node.end - 1:
                node.agg = node.start
                return docs.dims[node.start].copy(), [m[node.start] for m in docs.mets]
            d = docs.dims[node.start].copy()
            _, m = _aggregate_sorted(docs.dims[node.start:node.end], [x[node.start:node.end] for x in docs.mets], kinds, [])
            d[node.dim + 1:] = STAR_IN_FORWARD_INDEX
            node.agg = docs.append(d[None, :], [x[:1] for x in m])
            return d, [x[0] for x in m]
        star = [c for c in node.children if c.value == STAR]
        result = None
        if star:
            for c in node.children:
                r = aggregated(c)
                if c.value == STAR:
                    result = r
                    node.agg = c.agg
            return result
        rows = [aggregated(c) for c in node.children]
        d = rows[0][0].copy()
        mm = []
        for i, k in enumerate(kinds):
            col = np.array([r[1][i] for r in rows])
            mm.append(col.sum() if k in ("COUNT", "SUM") else col.max() if k == "MAX" else col.min())
        d[node.dim + 1:] = STAR_IN_FORWARD_INDEX
        node.agg = docs.append(d[None, :], [np.array([x]) for x in mm])
        return d, mm

    import sys
    sys.setrecursionlimit(10000)
    aggregated(root)

    # 4. serialize (BFS, children sorted by value)
    header = struct.pack("<QiiI", MAGIC, 1, 0, D)
    for i, name in enumerate(dimensions):
        b = name.encode("utf-8")
        header += struct.pack("<ii", i, len(b)) + b
    header += struct.pack("<i", num_nodes[0])
    header = header[:12] + struct.pack("<i", len(header)) + header[16:]
    rows = []
    queue = [root]
    cur = 0
    qi = 0
    while qi < len(queue):
        node = queue[qi]
        qi += 1
        if node.children is None:
            rows.append((node.dim, node.value, node.start, node.end, node.agg, -1, -1))
        else:
            ch = sorted(node.children, key=lambda c: c.value)
            first = cur + (len(queue) - qi) + 1
            rows.append((node.dim, node.value, node.start, node.end, node.agg, first, first + len(ch) - 1))
            queue.extend(ch)
        cur += 1
    assert len(rows) == num_nodes[0]
    tree = np.frombuffer(header + np.asarray(rows, dtype="<i4").tobytes(), dtype=np.uint8).copy()

    total = docs.n
    dim_ids = docs.dims[:total].copy()
    metrics = [m[:total].copy() for m in docs.mets]
    # star-tree docs as a segment: dimensions keep the base column's dictionary, bits and cardinality
    cols = []
    for j, dname in enumerate(dimensions):
        base = seg.column(dname)
        cols.append(sb.ColumnData(dname, base.data_type, True, base.bits, base.cardinality, False, base.dict_entry_bytes,
                                  sb.pack_fixed_bits(dim_ids[:, j], base.bits), base.dict, None, base.dict_values,
                                  dim_ids[:, j].copy()))
    st = StarTreeData(list(dimensions), list(function_pairs), total, tree, dim_ids, metrics, None, max_leaf_records)
    for i, m in enumerate(metrics):
        cols.append(sb.build_raw_column(st.metric_name(i), m))
    st.segment = sb.SegmentData(seg.name + "$startree", total, cols)
    return st
