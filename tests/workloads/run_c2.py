#!/usr/bin/env python
"""The c2 record of bench.py alone (BASELINE.json configs[1]: 8 x 100 M rows, 2-predicate range filter + SUM / COUNT), for
profiling the aggregation-only kernel:  python tests/workloads/run_c2.py [--steps N] [--selectivity 0.25]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--selectivity", type=float, default=0.25)
    ap.add_argument("--segments", type=int, default=8)
    ap.add_argument("--rows", type=int, default=100_000_000)
    args = ap.parse_args()
    from pinot_b200 import sql
    from pinot_b200.plan_maker import B200Context, B200PlanMaker, IndexSegment
    ctx = B200Context(0)
    pm = B200PlanMaker(ctx)
    segs = [IndexSegment.synthetic(ctx, f"r0s{s}", args.rows, bench.column_specs(0, s)) for s in range(args.segments)]
    q = sql.parse(bench.c2_query_text(args.selectivity))
    for _ in range(args.warmup):
        pm.execute_segments(segs, q)
    t0 = time.perf_counter()
    kms = []
    for _ in range(args.steps):
        blocks = pm.execute_segments(segs, q)
        kms.append(blocks[0].device_ms)
    el = time.perf_counter() - t0
    rows = args.segments * args.rows
    k = sum(kms) / len(kms)
    print(json.dumps({"workload": "c2", "query": bench.c2_query_text(args.selectivity), "ms_per_step": el / args.steps * 1e3,
                      "kernel_ms": k, "achieved_gbs": rows * bench.bytes_per_row() / (k * 1e-3) / 1e9,
                      "count": sum(int(b.longs[1][0]) for b in blocks)}))
    for s in segs:
        s.destroy()
    ctx.close()


if __name__ == "__main__":
    main()
