#!/usr/bin/env python
"""BASELINE.json configs[4] ("C5"): star-tree pre-aggregated lookup vs raw scan on the SAME segment, GROUP BY a high-cardinality
dimension, 1 GPU.

    python tests/workloads/run_c5.py --rows 100000000

The base segment is generated in HBM (pb200_synth_segment), its forward indexes are read back and the star-tree is built
on the host with the test-side builder (oracle/startree_builder.py: segment GENERATION is offline work in the reference
too -- MultipleTreesBuilder runs at segment creation -- and is not on the measured path).  Then both plans are timed through
the public call (pb200h_execute): useStarTree=false -> raw scan group-by; default -> StarTreeFilterOperator traversal on
the host + the same scan kernel over the pre-aggregated docs.  Results of the two plans must be identical.

Prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--high-card", type=int, default=100_000)
    ap.add_argument("--low-card", type=int, default=10)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--max-leaf-records", type=int, default=10_000)
    args = ap.parse_args()

    import numpy as np
    from gpu_util import assert_tables_equal, gpu_table
    from oracle import segment_builder as sb
    from oracle import startree_builder as stb
    from oracle.pinot_oracle import oracle as get_oracle
    from pinot_b200 import sql
    from pinot_b200.plan_maker import B200Context, B200PlanMaker, IndexSegment

    o = get_oracle()
    ctx = B200Context(0)
    pm = B200PlanMaker(ctx)
    n = args.rows
    specs = [{"name": "d_hi", "cardinality": args.high_card, "value_base": 0, "value_step": 1, "seed": 11},
             {"name": "d_lo", "cardinality": args.low_card, "value_base": 0, "value_step": 1, "seed": 12},
             {"name": "m", "cardinality": 1000, "value_base": 1, "value_step": 1, "seed": 13}]
    seg = IndexSegment.synthetic(ctx, "c5", n, specs)

    t0 = time.perf_counter()
    all_docs = np.arange(n, dtype=np.int32)
    cols = []
    for sp in specs:
        info = seg.column_info(sp["name"])
        fwd = seg.read_index(sp["name"], "fwd")
        vals = (sp["value_base"] + sp["value_step"] * np.arange(sp["cardinality"])).astype(np.int32)
        ids = o.read_dict_ids(fwd, n, info["bits"], all_docs)
        cols.append(sb.ColumnData(sp["name"], sb.INT, True, info["bits"], sp["cardinality"], False, 4, fwd,
                                  seg.read_index(sp["name"], "dict"), None, dict_values=vals, dict_ids=ids))
    del all_docs
    host = sb.SegmentData("c5", n, cols)
    st = stb.build_star_tree(host, ["d_hi", "d_lo"], [("COUNT", None), ("SUM", "m"), ("MAX", "m")],
                             max_leaf_records=args.max_leaf_records)
    build_s = time.perf_counter() - t0
    nd = len(st.dimensions)
    seg.attach_star_tree(st.tree, st.num_docs, st.dimensions, [st.segment.columns[j].fwd for j in range(nd)],
                         [(fn, col, st.segment.columns[nd + i].fwd) for i, (fn, col) in enumerate(st.function_pairs)])

    out = {"workload": "C5", "rows": n, "star_tree_docs": int(st.num_docs), "star_tree_build_s": build_s,
           "dimensions": st.dimensions, "high_card": args.high_card, "queries": []}
    for text in ["SELECT SUM(m), COUNT(*), MAX(m) FROM t GROUP BY d_hi",
                 "SELECT SUM(m), COUNT(*) FROM t WHERE d_lo = 3 GROUP BY d_hi",
                 "SELECT SUM(m), COUNT(*) FROM t WHERE d_hi BETWEEN 1000 AND 1999"]:
        res = {}
        tables = {}
        for plan, use in (("raw_scan", False), ("star_tree", True)):
            q = sql.parse(text, use_star_tree=use, num_groups_limit=2_000_000)
            for _ in range(args.warmup):
                b = pm.execute_segments([seg], q, views=True)[0]
            t0 = time.perf_counter()
            for _ in range(args.steps):
                b = pm.execute_segments([seg], q, views=True)[0]
            ms = (time.perf_counter() - t0) / args.steps * 1e3
            assert b.operator_kind == ("STAR_TREE" if use else ("GROUP_BY" if q.is_group_by else "AGGREGATION")), b.operator_kind
            tables[plan] = gpu_table(host, q, b)
            res[plan] = {"ms_per_query": ms, "scan_kernel_ms": b.device_ms, "docs_scanned": b.stats.num_docs_scanned,
                         "groups": b.num_groups}
        assert_tables_equal(q, tables["star_tree"], tables["raw_scan"], text)
        res["speedup"] = res["raw_scan"]["ms_per_query"] / res["star_tree"]["ms_per_query"]
        res["query"] = text
        res["identical"] = True
        out["queries"].append(res)
    print(json.dumps(out))
    b = None
    seg.destroy()
    ctx.close()


if __name__ == "__main__":
    main()
