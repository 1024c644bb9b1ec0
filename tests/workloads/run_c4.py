#!/usr/bin/env python
"""BASELINE.json configs[3] ("C4"): segments sharded across GPUs, filter + GROUP BY 2 dims + SUM/MAX, one NCCL reduce of
the per-GPU group tables.  Run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/workloads/run_c4.py --segments-per-gpu 8 --rows 50000000

Every rank generates its own segments (shared dictionaries: same cardinalities / value maps, different seeds), scans
them with the device-side combine, the dense tables are reduced to rank 0 over NCCL, rank 0 extracts the groups.  The
result is checked against the merge of the per-rank tables done in Python (``--check``), and the step is timed on the
device (max over ranks).  Prints one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

COLS = [("f", 10_000), ("g1", 1_000), ("g2", 100), ("m1", 100_000), ("m2", 65_536)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--segments-per-gpu", type=int, default=8)
    ap.add_argument("--rows", type=int, default=50_000_000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from pinot_b200 import sql
    from pinot_b200.distributed import execute_and_combine, init_comm
    from pinot_b200.plan_maker import B200Context, B200PlanMaker, IndexSegment

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = B200Context(local)
    pm = B200PlanMaker(ctx)
    if world > 1 and os.environ.get("PB200_TORCH_REDUCE", "0") != "1":
        init_comm(ctx, dist)
    segs = [IndexSegment.synthetic(ctx, f"r{rank}s{s}", args.rows,
                                   [{"name": n, "cardinality": c, "value_base": 1, "value_step": 3,
                                     "seed": 7919 * (rank * 1000 + s) + i} for i, (n, c) in enumerate(COLS)])
            for s in range(args.segments_per_gpu)]
    # f BETWEEN selects dictIds [1000, 2000) = 10 %
    q = sql.parse("SELECT SUM(m1), MAX(m2) FROM t WHERE f BETWEEN 3001 AND 5998 GROUP BY g1, g2",
                  num_groups_limit=1_000_000)

    from pinot_b200.distributed import DeviceBackend
    backend = DeviceBackend(pm, views=True)   # the root's results block aliases the pinned block the device extracted into

    def step():
        if world > 1:  # dense tables stay on the device, are reduced once over NCCL, rank 0 extracts the groups
            return execute_and_combine(backend, segs, q, dist, dst=0)
        return pm.execute_segments(segs, q, merge=True, views=True)[0]

    for _ in range(args.warmup):
        out = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([el], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t.item())

    ok = None
    if args.check:
        local_block = pm.execute_segments(segs, q, merge=True)[0]
        mine = {tuple(int(k) for k in local_block.keys[i]): (float(local_block.doubles[0][i]), float(local_block.doubles[1][i]), 0)
                for i in range(local_block.num_groups)}
        gathered = [None] * world
        if world > 1:
            dist.all_gather_object(gathered, mine)
        else:
            gathered = [mine]
        if rank == 0:
            want = {}
            for part in gathered:
                for k, (s, mx, c) in part.items():
                    if k in want:
                        want[k] = (want[k][0] + s, max(want[k][1], mx), want[k][2] + c)
                    else:
                        want[k] = (s, mx, c)
            got = {tuple(int(k) for k in out.keys[i]): (float(out.doubles[0][i]), float(out.doubles[1][i]), 0)
                   for i in range(out.num_groups)}
            ok = got == want
    if rank == 0:
        rows = world * args.segments_per_gpu * args.rows
        print(json.dumps({"workload": "C4", "n_gpus": world, "segments_per_gpu": args.segments_per_gpu, "rows_per_segment": args.rows,
                          "groups": out.num_groups, "ms_per_step": el / args.steps * 1e3, "rows_per_s": rows / (el / args.steps),
                          "kernel_ms_rank0": pm.last_device_ms, "check": ok}))
    out = None
    for s in segs:
        s.destroy()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
