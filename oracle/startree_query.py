"""Star-tree query evaluation with the oracle -- TEST INFRASTRUCTURE.

Restates the plan-level substitution of ``AggregationFunctionUtils.buildAggregationInfo`` /
``StarTreeUtils`` (``pinot-core/.../core/startree/StarTreeUtils.java``): a query fits a star-tree when every filter
predicate column and group-by column is a star-tree dimension, the filter is a conjunction of predicates, and every
aggregation maps to a function-column pair; then ``StarTreeFilterOperator`` (traversal + remaining predicates) feeds
the same aggregation operators, which read the pre-aggregated metric columns:

    COUNT(*) -> SUM over count__*      SUM(c) -> SUM over sum__c      MIN/MAX(c) -> MIN/MAX over min__c / max__c
    AVG(c)   -> (SUM over sum__c, SUM over count__*)
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from pinot_b200.query import Aggregation, Filter, Predicate, QueryContext

from . import segment_builder as sb


def matching_dict_ids(col: sb.ColumnData, p: Predicate) -> np.ndarray:
    """PredicateEvaluator.getMatchingDictIds on the base column's sorted dictionary."""
    vals = col.dict_values
    enc = (lambda v: v.encode() if isinstance(v, str) else v) if col.data_type == sb.STRING else (lambda v: v)
    card = len(vals)
    if p.type == "RANGE":
        lo = 0 if p.lower is None else int(np.searchsorted(vals, enc(p.lower), side="left" if p.lower_inclusive else "right"))
        hi = card if p.upper is None else int(np.searchsorted(vals, enc(p.upper), side="right" if p.upper_inclusive else "left"))
        return np.arange(lo, max(hi, lo), dtype=np.int32)
    ids = []
    for v in p.values:
        i = int(np.searchsorted(vals, enc(v)))
        if i < card and vals[i] == enc(v):
            ids.append(i)
    ids = np.unique(np.asarray(ids, dtype=np.int32))
    if p.type in ("EQ", "IN"):
        return ids
    return np.setdiff1d(np.arange(card, dtype=np.int32), ids)


def star_plan(seg: sb.SegmentData, st, q: QueryContext):
    """Returns (predicates per star dimension {dim: ids}, group-by dims, star aggregations, mapping) or None if unfit."""
    preds: Dict[int, np.ndarray] = {}
    leaves: List[Predicate] = []
    if q.filter is not None:
        if isinstance(q.filter, Predicate):
            leaves = [q.filter]
        elif q.filter.type == "AND" and all(isinstance(c, Predicate) for c in q.filter.children):
            leaves = list(q.filter.children)
        else:
            return None
    for p in leaves:
        if p.column not in st.dimensions:
            return None
        d = st.dimensions.index(p.column)
        ids = matching_dict_ids(seg.column(p.column), p)
        preds[d] = ids if d not in preds else np.intersect1d(preds[d], ids)
    if any(g not in st.dimensions for g in q.group_by):
        return None
    pairs = {(fn, col): i for i, (fn, col) in enumerate(st.function_pairs)}
    star_aggs: List[Aggregation] = []
    mapping = []  # per original aggregation: (kind, star agg indices)

    def use(fn, col):
        key = (fn, col)
        if key not in pairs:
            return None
        name = st.metric_name(pairs[key])
        a = Aggregation("SUM" if fn in ("COUNT", "SUM") else fn, name)
        star_aggs.append(a)
        return len(star_aggs) - 1

    for a in q.aggregations:
        if a.function == "COUNT":
            i = use("COUNT", None)
            idx = (i,)
        elif a.function == "AVG":
            idx = (use("SUM", a.column), use("COUNT", None))
        elif a.function in ("SUM", "MIN", "MAX"):
            idx = (use(a.function, a.column),)
        else:
            return None
        if any(i is None for i in idx):
            return None
        mapping.append((a.function, idx))
    return preds, [st.dimensions.index(g) for g in q.group_by], star_aggs, mapping, leaves


def execute_with_star_tree(oracle, seg: sb.SegmentData, st, q: QueryContext):
    """Returns {key values: [intermediates]} like tests/reduce_util.normalise, or None when the query does not fit."""
    plan = star_plan(seg, st, q)
    if plan is None:
        return None
    preds, gb_dims, star_aggs, mapping, leaves = plan
    docs, remaining = oracle.startree_traverse(st.tree, preds, gb_dims, st.num_docs)
    if docs is None:
        return {} if q.group_by else {(): [0 if fn == "COUNT" else (0.0, 0) if fn == "AVG" else
                                            (float("inf") if fn == "MIN" else float("-inf") if fn == "MAX" else 0.0)
                                            for fn, _ in mapping]}
    rem = [p for p in leaves if st.dimensions.index(p.column) in remaining]
    flt = None if not rem else rem[0] if len(rem) == 1 else Filter("AND", rem)
    sq = QueryContext(aggregations=star_aggs, filter=flt, group_by=list(q.group_by), num_groups_limit=q.num_groups_limit,
                      max_initial_result_holder_capacity=q.max_initial_result_holder_capacity)
    r = oracle.execute(st.segment, sq, doc_ids=docs)
    rows = 1 if r.num_groups < 0 else r.num_groups
    out = {}
    for g in range(rows):
        key = () if r.num_groups < 0 else tuple(st.segment.value_of(c, int(r.keys[g, j])) for j, c in enumerate(q.group_by))
        vals = []
        for fn, idx in mapping:
            if fn == "COUNT":
                vals.append(int(r.doubles[idx[0]][g]))
            elif fn == "AVG":
                vals.append((float(r.doubles[idx[0]][g]), int(r.doubles[idx[1]][g])))
            else:
                vals.append(float(r.doubles[idx[0]][g]))
        out[key] = vals
    return out
