#!/usr/bin/env python
"""Where a step's host time goes on ONE GPU: the library's own phase clock (pb200_last_phases) + the Python side
(marshalling + pb200h_execute call vs reading the result into numpy), for the bench headline, C4 and a 1 M-group query."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def profile(pm, segs, q, merge, steps=30):
    import ctypes as C
    from pinot_b200 import _lib
    from pinot_b200 import plan_maker as P
    ctx = pm.ctx
    for _ in range(3):
        pm.execute_segments(segs, q, merge=merge)
    t_all = time.perf_counter()
    for _ in range(steps):
        pm.execute_segments(segs, q, merge=merge)
    t_all = (time.perf_counter() - t_all) / steps * 1e3
    hq, _keep = P._marshal_query(q, merge, 0, False, 0)
    n = len(segs)
    sh = (C.c_void_p * n)(*[s.handle for s in segs])
    nres = 1 if merge else n
    acc = {"call": 0.0, "read": 0.0}
    phases = {}
    for _ in range(steps):
        res = (C.c_void_p * nres)(); kinds = (C.c_int32 * n)()
        t0 = time.perf_counter()
        _lib.check(ctx.lib.pb200h_execute(ctx.handle, C.byref(hq), sh, n, res, kinds))
        t1 = time.perf_counter()
        for k, v in ctx.last_phases().items():
            phases[k] = phases.get(k, 0.0) + v
        t1b = time.perf_counter()
        blocks = [P._read_result(ctx, C.c_void_p(res[i]), q, kinds[0], False) for i in range(nres)]
        t2 = time.perf_counter()
        acc["call"] += t1 - t0
        acc["read"] += t2 - t1b
    return {"step_ms": round(t_all, 4), "execute_call_ms": round(acc["call"] / steps * 1e3, 4), "python_read_ms": round(acc["read"] / steps * 1e3, 4),
            "library_phases_ms": {k: round(v / steps, 4) for k, v in phases.items()}, "kernel_ms": round(blocks[0].device_ms, 4),
            "groups": sum(max(b.num_groups, 0) for b in blocks)}


def main():
    from pinot_b200 import sql
    from pinot_b200.plan_maker import B200Context, B200PlanMaker, IndexSegment
    ctx = B200Context(0)
    pm = B200PlanMaker(ctx)
    out = {}
    segs = [IndexSegment.synthetic(ctx, f"s{s}", 100_000_000, bench.column_specs(0, s)) for s in range(8)]
    out["headline (10 000 groups, merged)"] = profile(pm, segs, sql.parse(bench.groupby_query_text(0.10)), True)
    out["headline, per-segment results"] = profile(pm, segs, sql.parse(bench.groupby_query_text(0.10)), False)
    for s in segs:
        s.destroy()
    cols = [("f", 10_000), ("g1", 1_000), ("g2", 100), ("m1", 100_000), ("m2", 65_536), ("g3", 1_000)]
    segs = [IndexSegment.synthetic(ctx, f"c4s{s}", 50_000_000,
                                   [{"name": n, "cardinality": c, "value_base": 1, "value_step": 3, "seed": 7919 * s + i} for i, (n, c) in enumerate(cols)])
            for s in range(8)]
    out["C4 (100 000 groups, merged)"] = profile(pm, segs, sql.parse(
        "SELECT SUM(m1), MAX(m2) FROM t WHERE f BETWEEN 3001 AND 5998 GROUP BY g1, g2", num_groups_limit=1_000_000), True)
    out["1 M groups (g1, g3), SUM + COUNT, merged"] = profile(pm, segs, sql.parse(
        "SELECT SUM(m1), COUNT(*) FROM t WHERE f BETWEEN 3001 AND 5998 GROUP BY g1, g3", num_groups_limit=2_000_000), True, steps=10)
    for s in segs:
        s.destroy()
    ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
