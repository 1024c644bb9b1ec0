#!/usr/bin/env python
"""BASELINE.json configs[2] ("C3"): inverted-index bitmap filter (3 EQ predicates AND-ed) -> GROUP BY dim (card 10 000)
SUM, 8 segments on 1 GPU.  Also C4-style group-by with a range filter (``--mode range``).

    python tests/workloads/run_c3.py --segments 8 --rows 100000000 [--check-rows 2000000]

Prints one JSON line: rows/s, device time of the scan kernel, algorithmic bytes (full-column figure of SURVEY section 8d:
(bits(g) + bits(m))/8 + 1/8 per doc mask read) and the achieved fraction of the measured HBM peak.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

COLS = [("d1", 10, True), ("d2", 20, True), ("d3", 50, True), ("g", 10_000, False), ("m", 100_000, False),
        ("f", 10_000, False), ("g2", 100, False)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--segments", type=int, default=8)
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", default="bitmap", choices=["bitmap", "range", "range2", "range3"])
    ap.add_argument("--check-rows", type=int, default=0, help="also verify against the oracle on a small segment")
    ap.add_argument("--query", default="", help="time this query instead of the mode's (columns d1 d2 d3 g m f g2); --bits = touched bits per row")
    ap.add_argument("--bits", type=int, default=0)
    ap.add_argument("--variants", default="", help="space separated name:knob=value,knob=value tuning variants "
                    "(pb200_tuning_set) timed on the same resident segments, one JSON line each")
    args = ap.parse_args()

    import numpy as np
    from pinot_b200 import sql
    from pinot_b200.plan_maker import B200Context, B200PlanMaker, IndexSegment

    ctx = B200Context(0)
    pm = B200PlanMaker(ctx)

    def specs(s, inverted=True):
        return [{"name": n, "cardinality": c, "value_base": 2, "value_step": 3, "inverted": inv and inverted,
                 "seed": 31337 * (s + 1) + i} for i, (n, c, inv) in enumerate(COLS)]

    if args.mode == "bitmap":
        text = "SELECT SUM(m), COUNT(*) FROM t WHERE d1 = 11 AND d2 = 17 AND d3 = 77 GROUP BY g"   # dictIds 3, 5, 25
        bits = 14 + 17 + 3  # g, m, three doc masks (1 bit each)
    elif args.mode == "range":
        text = "SELECT SUM(m), COUNT(*) FROM t WHERE f BETWEEN 3002 AND 5999 GROUP BY g"             # 10 %
        bits = 14 + 14 + 17
    elif args.mode == "range3":
        text = "SELECT SUM(m), COUNT(*) FROM t WHERE f BETWEEN 3002 AND 5999 GROUP BY g2"                 # 100 groups
        bits = 14 + 7 + 17
    else:
        text = "SELECT SUM(m), MAX(f), COUNT(*) FROM t WHERE f BETWEEN 3002 AND 5999 GROUP BY g, g2"  # 1M-group key space
        bits = 14 + 14 + 7 + 17
    if args.query:
        text, bits = args.query, args.bits
    q = sql.parse(text, num_groups_limit=2_000_000)

    if args.check_rows:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from gpu_util import assert_tables_equal, gpu_table, oracle_table
        from oracle import segment_builder as sb
        from oracle.pinot_oracle import oracle as get_oracle
        o = get_oracle()
        n = args.check_rows
        seg = IndexSegment.synthetic(ctx, "chk", n, specs(0))
        cols = []
        for sp in specs(0):
            info = seg.column_info(sp["name"])
            vals = (sp["value_base"] + sp["value_step"] * np.arange(sp["cardinality"])).astype(np.int32)
            inv = seg.read_index(sp["name"], "inv") if sp["inverted"] else None
            cols.append(sb.ColumnData(sp["name"], sb.INT, True, info["bits"], sp["cardinality"], False, 4,
                                      seg.read_index(sp["name"], "fwd"), seg.read_index(sp["name"], "dict"), inv, dict_values=vals))
        host = sb.SegmentData("chk", n, cols)
        assert_tables_equal(q, gpu_table(host, q, pm.make_segment_plan_node(seg, q).run().next_block()),
                            oracle_table(host, q, o.execute(host, q)), "check")
        seg.destroy()

    t0 = time.perf_counter()
    segs = [IndexSegment.synthetic(ctx, f"s{s}", args.rows, specs(s, inverted=args.mode == "bitmap")) for s in range(args.segments)]
    gen_s = time.perf_counter() - t0
    defaults = {"warps": 6, "ctas_per_sm": 2, "stages": 0, "grid": 0, "sparse_max": 4, "sparse_max_agg": -1, "smem_groups": 1,
                "smem_groups_max": 2048, "smem_copies": 0, "gb_defer": 1, "skip": 1, "always_count": 0, "pack_count": 1, "pack_shift": 0, "table_stride": 0}
    rows = args.segments * args.rows
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    reference_result = None
    for variant in (args.variants.split() or ["default:"]):
        vname, _, knobs = variant.partition(":")
        for k, v in defaults.items():
            ctx.set_tuning(k, v)
        for kv in filter(None, knobs.split(",")):
            k, _, v = kv.partition("=")
            ctx.set_tuning(k, int(v))
        for _ in range(args.warmup):
            blocks = pm.execute_segments(segs, q)
        t0 = time.perf_counter()
        kms = []
        for _ in range(args.steps):
            blocks = pm.execute_segments(segs, q)
            kms.append(blocks[0].device_ms)
        el = time.perf_counter() - t0
        # every variant must return the same tables
        digest = [(b.num_groups, float(sum(d.sum() for d in b.doubles)), int(b.longs[-1].sum()), b.stats.num_docs_scanned) for b in blocks]
        if reference_result is None:
            reference_result = digest
        k = sum(kms) / len(kms)
        print(json.dumps({"workload": f"C3/{args.mode}", "variant": vname, "knobs": knobs, "query": text, "segments": args.segments,
                          "rows_per_segment": args.rows, "ms_per_step": el / args.steps * 1e3, "rows_per_s": rows / (el / args.steps),
                          "scan_kernel_ms": k, "algorithmic_bits_per_row": bits, "achieved_gbs": rows * bits / 8 / (k * 1e-3) / 1e9,
                          "frac_of_peak": rows * bits / 8 / (k * 1e-3) / 1e9 / peak, "groups": [b.num_groups for b in blocks][:3],
                          "matched": sum(b.stats.num_docs_scanned for b in blocks), "generation_s": gen_s,
                          "same_as_first_variant": digest == reference_result, "checked": bool(args.check_rows)}), flush=True)
    for s in segs:
        s.destroy()
    ctx.close()


if __name__ == "__main__":
    main()
