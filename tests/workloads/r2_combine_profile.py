#!/usr/bin/env python
"""Where the cross-GPU combine's time goes (torchrun, N ranks): phases of pinot_b200.distributed.execute_and_combine on the
bench headline, wall-clock per phase on rank 0 with a device sync after each (so the sum exceeds the pipelined total)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from pinot_b200 import sql
    from pinot_b200.distributed import DeviceBackend, device_buffers, execute_and_combine, global_domain, reduce_buffers
    from pinot_b200.plan_maker import B200Context, B200PlanMaker, IndexSegment
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = B200Context(local)
    pm = B200PlanMaker(ctx)
    rows = 100_000_000
    segs = [IndexSegment.synthetic(ctx, f"r{rank}s{s}", rows, bench.column_specs(rank, s)) for s in range(8)]
    dom = global_domain(ctx, segs, ["c3", "c5"], dist)
    q = sql.parse(bench.groupby_query_text(0.10))
    be = DeviceBackend(pm)
    bound = 8 * rows * world
    for _ in range(5):
        execute_and_combine(pm, segs, q, dist, 0, bound)
    T = {k: 0.0 for k in ("execute", "flag_allreduce", "buffers", "reduce", "sync_item", "finish", "total_pipelined")}
    steps = 50

    def tick():
        torch.cuda.synchronize(local)
        return time.perf_counter()
    for _ in range(steps):
        dist.barrier(); t0 = tick()
        block = be.execute(segs, q, world, bound, False); t1 = tick()
        flag = be.flag_tensor(1 if block.carrier_unsafe else 0)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX); t2 = tick()
        bufs = be.buffers(block); t3 = tick()
        reduce_buffers(bufs, dist, 0); t4 = tick()
        be.synchronize(); f = int(flag.item()); t5 = tick()
        out = be.finish(block, q, rank == 0); t6 = tick()
        for k, a, b in (("execute", t0, t1), ("flag_allreduce", t1, t2), ("buffers", t2, t3), ("reduce", t3, t4), ("sync_item", t4, t5), ("finish", t5, t6)):
            T[k] += b - a
    dist.barrier(); t0 = tick()
    for _ in range(steps):
        execute_and_combine(pm, segs, q, dist, 0, bound)
    dist.barrier(); T["total_pipelined"] = tick() - t0
    if rank == 0:
        print(json.dumps({"world": world, "ms_per_step": {k: round(v / steps * 1e3, 4) for k, v in T.items()},
                          "table_bytes": {k: (0 if v is None else v.numel() * v.element_size()) for k, v in bufs.items()},
                          "kernel_ms": pm.last_device_ms, "carrier": block.count_carrier}))
    for s in segs:
        s.destroy()
    dom.release()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
