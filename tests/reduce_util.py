"""Test-side restatement of what sits ABOVE the path: the per-server combine and the broker reduce.

GroupByCombineOperator merges per-segment blocks by group-key VALUES with AggregationFunction.merge
(pinot-core/.../operator/combine/GroupByCombineOperator.java:126-156, data/table/IndexedTable.java:101-136); the broker
then extracts final results, orders and trims (query/reduce/GroupByDataTableReducer.java).  Both stay Java in the real
integration; tests need them only to compare against the reference's inter-segment goldens.

A segment result is normalised to  {key_values_tuple: [intermediate per aggregation]}  with intermediates
COUNT -> int, SUM/MIN/MAX -> float, AVG -> (sum, count), DISTINCTCOUNT -> frozenset of VALUES.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple


def merge_intermediate(fn: str, a, b):
    if fn in ("COUNT", "SUM"):
        return a + b
    if fn == "MIN":
        return min(a, b)
    if fn == "MAX":
        return max(a, b)
    if fn == "AVG":
        return (a[0] + b[0], a[1] + b[1])
    if fn == "DISTINCTCOUNT":
        return a | b
    raise ValueError(fn)


def final_value(fn: str, v):
    if fn == "AVG":
        return v[0] / v[1] if v[1] else float("-inf")
    if fn == "DISTINCTCOUNT":
        return len(v)
    return v


def combine(functions: Sequence[str], blocks: Sequence[Dict[tuple, list]]) -> Dict[tuple, list]:
    out: Dict[tuple, list] = {}
    for b in blocks:
        for key, vals in b.items():
            if key not in out:
                out[key] = list(vals)
            else:
                out[key] = [merge_intermediate(f, x, y) for f, x, y in zip(functions, out[key], vals)]
    return out


def reduce_rows(functions: Sequence[str], table: Dict[tuple, list]) -> List[Tuple[tuple, list]]:
    return [(k, [final_value(f, v) for f, v in zip(functions, vals)]) for k, vals in table.items()]


def normalise(seg, query, num_groups, keys, doubles, longs, distinct) -> Dict[tuple, list]:
    """Turns dictId-space results (oracle or GPU) of one segment into the value-space dict described above.

    `seg` is an oracle.segment_builder.SegmentData (dictionary values), `keys` [G, k] dictIds,
    `doubles`/`longs` per aggregation arrays, `distinct` {(agg, group): dictIds}.
    """
    rows = 1 if num_groups < 0 else num_groups
    out = {}
    for g in range(rows):
        key = () if num_groups < 0 else tuple(seg.value_of(c, int(keys[g, j])) for j, c in enumerate(query.group_by))
        vals = []
        for a, agg in enumerate(query.aggregations):
            fn = agg.function
            if fn == "COUNT":
                vals.append(int(longs[a][g]))
            elif fn == "AVG":
                vals.append((float(doubles[a][g]), int(longs[a][g])))
            elif fn == "DISTINCTCOUNT":
                vals.append(frozenset(seg.value_of(agg.column, int(d)) for d in distinct[(a, g)]))
            else:
                vals.append(float(doubles[a][g]))
        out[key] = vals
    return out
