"""Host-side query model: the slice of Pinot's ``QueryContext`` the scan -> filter -> group-by path consumes.

Mirrors (names and meaning, not code) ``pinot-core/src/main/java/org/apache/pinot/core/query/request/context/
QueryContext.java``: a filter tree (``FilterContext`` with AND / OR / NOT / PREDICATE nodes,
``pinot-common/.../request/context/FilterContext.java``), group-by expressions (identifiers only on this path),
aggregation functions and the two server options that shape group-by execution (``numGroupsLimit`` and
``maxInitialResultHolderCapacity``: ``core/plan/maker/InstancePlanMakerImplV2.java:68-91``).

Predicates are in VALUE space here (as the broker sends them); turning them into dictId space is the plan maker's job
(``pinot_b200/plan_maker.py``), exactly where the reference runs its ``PredicateEvaluator``s.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Union

# InstancePlanMakerImplV2 defaults (core/plan/maker/InstancePlanMakerImplV2.java:68-82)
DEFAULT_MAX_INITIAL_RESULT_HOLDER_CAPACITY = 10_000
DEFAULT_NUM_GROUPS_LIMIT = 100_000

Literal = Union[int, float, str]


@dataclass
class Predicate:
    """Leaf of the filter tree: ``Predicate.Type`` EQ / NOT_EQ / IN / NOT_IN / RANGE on an identifier."""
    type: str  # "EQ" | "NEQ" | "IN" | "NOT_IN" | "RANGE"
    column: str
    values: List[Literal] = field(default_factory=list)  # EQ/NEQ: [v]; IN/NOT_IN: [v...]
    lower: Optional[Literal] = None  # RANGE; None = unbounded ("*")
    upper: Optional[Literal] = None
    lower_inclusive: bool = True
    upper_inclusive: bool = True


@dataclass
class Filter:
    """AND / OR / NOT node (``FilterContext.Type``)."""
    type: str  # "AND" | "OR" | "NOT"
    children: List[Union["Filter", Predicate]] = field(default_factory=list)


FilterNode = Union[Filter, Predicate]


@dataclass
class Aggregation:
    function: str  # COUNT | SUM | MIN | MAX | AVG | DISTINCTCOUNT
    column: Optional[str] = None  # None for COUNT(*)
    # FILTER (WHERE ...) clause of this function (QueryContext.getFilteredAggregationFunctions: Pair<function, FilterContext>);
    # evaluated together with the query's own filter (AggregationFunctionUtils.buildFilteredAggregationInfos)
    filter: Optional["FilterNode"] = None

    def __str__(self):
        return f"{self.function.lower()}({self.column or '*'})"


@dataclass
class QueryContext:
    aggregations: List[Aggregation]
    filter: Optional[FilterNode] = None
    group_by: List[str] = field(default_factory=list)
    table: str = "testTable"
    limit: int = 10
    order_by: List[tuple] = field(default_factory=list)  # (expression string, ascending)
    num_groups_limit: int = DEFAULT_NUM_GROUPS_LIMIT
    max_initial_result_holder_capacity: int = DEFAULT_MAX_INITIAL_RESULT_HOLDER_CAPACITY
    and_scan_reordering: bool = False
    use_star_tree: bool = True  # query option useStarTree

    @property
    def is_group_by(self) -> bool:
        return len(self.group_by) > 0


# ----------------------------------------------------------------------------------------------------------------------
# Broker-side filter optimizers that shape what the server receives
# (core/query/optimizer/filter/{FlattenAndOr,MergeEqIn,MergeRange}FilterOptimizer.java)
# ----------------------------------------------------------------------------------------------------------------------
def flatten(node: Optional[FilterNode]) -> Optional[FilterNode]:
    if node is None or isinstance(node, Predicate):
        return node
    kids = [flatten(c) for c in node.children]
    if node.type in ("AND", "OR"):
        flat = []
        for k in kids:
            if isinstance(k, Filter) and k.type == node.type:
                flat.extend(k.children)
            else:
                flat.append(k)
        kids = flat
        if len(kids) == 1:
            return kids[0]
    return Filter(node.type, kids)


def merge_eq_in(node: Optional[FilterNode]) -> Optional[FilterNode]:
    """OR of EQ/IN on the same column -> one IN (MergeEqInFilterOptimizer)."""
    if node is None or isinstance(node, Predicate):
        return node
    kids = [merge_eq_in(c) for c in node.children]
    if node.type != "OR":
        return Filter(node.type, kids)
    by_col = {}
    rest = []
    for k in kids:
        if isinstance(k, Predicate) and k.type in ("EQ", "IN"):
            by_col.setdefault(k.column, []).append(k)
        else:
            rest.append(k)
    merged = []
    for col, preds in by_col.items():
        if len(preds) == 1:
            merged.append(preds[0])
        else:
            vals = []
            for p in preds:
                for v in p.values:
                    if v not in vals:
                        vals.append(v)
            merged.append(Predicate("EQ", col, vals) if len(vals) == 1 else Predicate("IN", col, vals))
    out = merged + rest
    return out[0] if len(out) == 1 else Filter("OR", out)


def merge_range(node: Optional[FilterNode]) -> Optional[FilterNode]:
    """AND of several RANGE predicates on the same column -> their intersection (MergeRangeFilterOptimizer)."""
    if node is None or isinstance(node, Predicate):
        return node
    kids = [merge_range(c) for c in node.children]
    if node.type != "AND":
        return Filter(node.type, kids)
    ranges = {}
    others = []
    recreate = False
    for k in kids:
        if isinstance(k, Predicate) and k.type == "RANGE":
            cur = ranges.get(k.column)
            if cur is None:
                ranges[k.column] = Predicate("RANGE", k.column, [], k.lower, k.upper, k.lower_inclusive, k.upper_inclusive)
            else:
                recreate = True
                if k.lower is not None and (cur.lower is None or k.lower > cur.lower or
                                            (k.lower == cur.lower and not k.lower_inclusive)):
                    cur.lower, cur.lower_inclusive = k.lower, k.lower_inclusive
                if k.upper is not None and (cur.upper is None or k.upper < cur.upper or
                                            (k.upper == cur.upper and not k.upper_inclusive)):
                    cur.upper, cur.upper_inclusive = k.upper, k.upper_inclusive
        else:
            others.append(k)
    if not recreate:
        return Filter("AND", kids)
    out = others + list(ranges.values())
    return out[0] if len(out) == 1 else Filter("AND", out)


def optimize_filter(node: Optional[FilterNode]) -> Optional[FilterNode]:
    return merge_range(merge_eq_in(flatten(node)))


def filter_columns(node: Optional[FilterNode]) -> List[str]:
    out: List[str] = []

    def walk(n):
        if n is None:
            return
        if isinstance(n, Predicate):
            if n.column not in out:
                out.append(n.column)
        else:
            for c in n.children:
                walk(c)

    walk(node)
    return out


def postfix(node: Optional[FilterNode]) -> Sequence[FilterNode]:
    """Children-before-parent order (what both the C-ABI and the oracle consume)."""
    out = []

    def walk(n):
        if isinstance(n, Filter):
            for c in n.children:
                walk(c)
        out.append(n)

    if node is not None:
        walk(node)
    return out
