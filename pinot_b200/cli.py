"""Command line face of the path: load Pinot segment directories into HBM and run one aggregation / group-by query on them.

    python -m pinot_b200.cli --segment-dir /data/myTable_0 --segment-dir /data/myTable_1 \\
        "SELECT SUM(m), COUNT(*) FROM myTable WHERE d > 10 GROUP BY k LIMIT 20"

What happens: ``pb200h_segment_load_dir`` per directory (v1 file-per-index or v3 columns.psf + index_map, raw columns decoded
and dictionary-encoded at load, star-trees attached), ``pb200h_execute`` over all segments in one submission, then -- what the
reference's combine + broker reduce do, in a few lines of Python -- the per-segment results blocks are merged by key VALUES
(``AggregationFunction.merge``), final results extracted (AVG = sum / count, DISTINCTCOUNT = set size) and the first LIMIT rows
printed.  With ``--merge-on-device`` the segments are first bound to one dictionary domain and combined on the GPU.

Needs a GPU: the product has no CPU fallback.  Queries outside the accelerated set exit with the library's message (a Pinot
server would run the stock operator for them)."""
from __future__ import annotations

import argparse
import json
import sys
from typing import Dict, List


def merge_blocks(query, segments, blocks) -> Dict[tuple, list]:
    """{key values: [intermediate per aggregation]} over all blocks: COUNT -> int, SUM / MIN / MAX -> float,
    AVG -> (sum, count), DISTINCTCOUNT -> set of values (GroupByCombineOperator / AggregationFunction.merge)."""
    out: Dict[tuple, list] = {}
    for seg, b in zip(segments, blocks):
        rows = 1 if b.num_groups < 0 else b.num_groups
        for g in range(rows):
            key = () if b.num_groups < 0 else tuple(seg.dictionary_value(c, int(b.keys[g, j])) for j, c in enumerate(query.group_by))
            vals = []
            for a, agg in enumerate(query.aggregations):
                fn = agg.function
                if fn == "COUNT":
                    vals.append(int(b.longs[a][g]))
                elif fn == "AVG":
                    vals.append((float(b.doubles[a][g]), int(b.longs[a][g])))
                elif fn == "DISTINCTCOUNT":
                    vals.append({seg.dictionary_value(agg.column, int(d)) for d in b.distinct[(a, g)]})
                else:
                    vals.append(float(b.doubles[a][g]))
            cur = out.get(key)
            if cur is None:
                out[key] = vals
                continue
            for a, agg in enumerate(query.aggregations):
                fn = agg.function
                if fn in ("COUNT", "SUM"):
                    cur[a] += vals[a]
                elif fn == "MIN":
                    cur[a] = min(cur[a], vals[a])
                elif fn == "MAX":
                    cur[a] = max(cur[a], vals[a])
                elif fn == "AVG":
                    cur[a] = (cur[a][0] + vals[a][0], cur[a][1] + vals[a][1])
                else:
                    cur[a] |= vals[a]
    return out


def final_rows(query, table: Dict[tuple, list]) -> List[list]:
    """Broker-side extraction of final results + LIMIT (no ORDER BY: Pinot returns an arbitrary LIMIT-sized subset; here the
    rows are sorted by key to make the output stable)."""
    rows = []
    for key in sorted(table, key=lambda k: tuple(str(x) for x in k)):
        vals = []
        for agg, v in zip(query.aggregations, table[key]):
            if agg.function == "AVG":
                vals.append(v[0] / v[1] if v[1] else float("-inf"))
            elif agg.function == "DISTINCTCOUNT":
                vals.append(len(v))
            else:
                vals.append(v)
        rows.append(list(key) + vals)
    return rows[: query.limit]


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m pinot_b200.cli", description=__doc__.split("\n\n")[0])
    ap.add_argument("sql")
    ap.add_argument("--segment-dir", action="append", required=True, help="a Pinot segment directory (repeatable)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--merge-on-device", action="store_true", help="bind the segments to one dictionary domain and combine on the GPU")
    ap.add_argument("--num-groups-limit", type=int, default=100_000)
    ap.add_argument("--json", action="store_true", help="print one JSON object instead of a table")
    args = ap.parse_args(argv)

    from . import sql
    from ._lib import Pb200Error
    from .plan_maker import B200Context, B200PlanMaker, DictionaryDomain, IndexSegment

    query = sql.parse(args.sql, num_groups_limit=args.num_groups_limit)
    ctx = B200Context(args.device)
    segments, domain = [], None
    try:
        for d in args.segment_dir:
            segments.append(IndexSegment.load(ctx, d))
        pm = B200PlanMaker(ctx)
        try:
            if args.merge_on_device and len(segments) > 1:
                cols = list(dict.fromkeys(list(query.group_by) + [a.column for a in query.aggregations
                                                                  if a.column and a.function in ("MIN", "MAX", "DISTINCTCOUNT")]))
                if cols:
                    domain = DictionaryDomain.build(ctx, segments, cols)
                    for s in segments:
                        s.bind_domain(domain)
                blocks = pm.execute_segments(segments, query, merge=True)
                table = merge_blocks(query, segments[:1], blocks)
            else:
                blocks = pm.execute_segments(segments, query)
                table = merge_blocks(query, segments, blocks)
        except Pb200Error as e:
            sys.stderr.write(f"{e}\n")
            return 2
        rows = final_rows(query, table)
        names = list(query.group_by) + [str(a) for a in query.aggregations]
        stats = {"numDocsScanned": sum(b.stats.num_docs_scanned for b in blocks), "totalDocs": sum(s.num_docs for s in segments),
                 "numSegments": len(segments), "deviceMs": round(sum(b.device_ms for b in blocks[:1]), 4),
                 "operators": sorted({b.operator_kind for b in blocks})}
        if args.json:
            print(json.dumps({"columns": names, "rows": rows, "stats": stats}, default=str))
        else:
            print("\t".join(names))
            for r in rows:
                print("\t".join(str(x) for x in r))
            print(f"-- {json.dumps(stats)}", file=sys.stderr)
        return 0
    finally:
        blocks = None
        for s in segments:
            s.destroy()
        if domain is not None:
            domain.release()
        ctx.close()


if __name__ == "__main__":
    sys.exit(main())
