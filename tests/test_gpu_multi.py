"""-m gpu, needs >= 2 GPUs (skipped otherwise; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
one process per GPU over NCCL, segments with DIFFERENT per-segment dictionaries on every rank.

  * pinot_b200.distributed.global_domain: the ranks' dictionary unions are all-gathered, every rank builds the same
    table-wide dictionaries and re-encodes its segments into them;
  * execute_and_combine: per-GPU device-side combine, ONE reduce of the group tables to rank 0 (count carrier incl. the
    collective fallback when a packed sum field cannot be proven safe), extraction on rank 0;
  * the result must equal the oracle-side merge BY VALUE (tests/reduce_util.combine == GroupByCombineOperator /
    AggregationFunction.merge) of every segment of every rank, each run through the CPU oracle with its own dictionaries.
"""
import os
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

QUERIES = [
    "SELECT SUM(v), COUNT(*) FROM t WHERE v > -400 GROUP BY k, j",        # > 2048 keys: global tables, count carried in SUM(v)
    "SELECT SUM(v), COUNT(*) FROM t WHERE v > -400 GROUP BY k",           # small key space: CTA-private shared-memory tables
    "SELECT AVG(v), MAX(l), MIN(k) FROM t WHERE j != 2 GROUP BY k, j",      # carrier + MIN / MAX ids in domain space
    "SELECT SUM(f), MAX(v) FROM t GROUP BY j",                              # float sums, shared-memory sized key space
    "SELECT SUM(v) FROM t WHERE l > 5000000000 GROUP BY k",                 # count only as the group-exists marker
    "SELECT COUNT(*), SUM(v), MIN(l), MAX(f) FROM t WHERE v < 400",         # aggregation only
    "SELECT DISTINCTCOUNT(k), DISTINCTCOUNT(l), COUNT(*) FROM t WHERE v > 0",                 # id sets: merged on the host by set union
    "SELECT DISTINCTCOUNT(v), SUM(v), MAX(l) FROM t WHERE j < 6 GROUP BY j",
    # key space 600+ x 7 x 30 x 2000 > 2^24: per-GPU HASH tables, not element-wise reducible -> merged by key on the host
    "SELECT SUM(v), COUNT(*), MAX(f) FROM t WHERE v > 900 GROUP BY k, j, l, v",
]


def _worker(rank, world, port, out_dir, pack_shift, native):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gpu_util import oracle_table, to_device
    from oracle.pinot_oracle import oracle as get_oracle
    from pinot_b200 import sql
    from pinot_b200.distributed import execute_and_combine, global_domain, init_comm
    from pinot_b200.plan_maker import B200Context, B200PlanMaker
    o = get_oracle()
    rng = np.random.default_rng(77 + 13 * rank)
    segs = []
    for i, n in enumerate([30_000, 8193 + rank, 50_001]):   # per-segment dictionaries: different subsets everywhere
        pk = rng.choice(600, size=90 + 11 * i + 5 * rank, replace=False).astype(np.int32) * 3 - 200
        pl = rng.choice(10_000, size=30, replace=False).astype(np.int64) * 1_000_003
        pf = (rng.choice(300, size=25 + i, replace=False) / 8.0).astype(np.float64)
        pv = rng.choice(2000, size=300, replace=False).astype(np.int32) - 1000
        segs.append(o.build_segment(f"r{rank}s{i}", {
            "k": pk[rng.integers(0, len(pk), size=n)], "j": rng.integers(0, 7, size=n).astype(np.int32),
            "l": pl[rng.integers(0, len(pl), size=n)], "f": pf[rng.integers(0, len(pf), size=n)],
            "v": pv[rng.integers(0, len(pv), size=n)]}))
    ctx = B200Context(rank)
    pm = B200PlanMaker(ctx)
    if native:
        init_comm(ctx, dist)   # the reduce runs inside libpinot_b200.so (pb200_result_combine), torch only ships the NCCL id
    if pack_shift:
        ctx.set_tuning("pack_shift", pack_shift)
    devs = [to_device(ctx, s) for s in segs]
    dom = global_domain(ctx, devs, ["k", "j", "l", "f", "v"], dist)
    report = {}
    for text in QUERIES:
        q = sql.parse(text, num_groups_limit=1_000_000)
        got = execute_and_combine(pm, devs, q, dist, dst=0)
        mine = [oracle_table(s, q, o.execute(s, q)) for s in segs]
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        if rank == 0:
            rows = 1 if got.num_groups < 0 else got.num_groups
            table = {}
            for g in range(rows):
                key = () if got.num_groups < 0 else tuple(devs[0].dictionary_value(c, int(got.keys[g, j])) for j, c in enumerate(q.group_by))
                vals = []
                for a, agg in enumerate(q.aggregations):
                    vals.append(int(got.longs[a][g]) if agg.function == "COUNT" else
                                (float(got.doubles[a][g]), int(got.longs[a][g])) if agg.function == "AVG" else
                                frozenset(devs[0].dictionary_value(agg.column, int(d)) for d in got.distinct[(a, g)])
                                if agg.function == "DISTINCTCOUNT" else float(got.doubles[a][g]))
                table[key] = vals
            report[text] = (table, [t for part in everyone for t in part], got.count_carrier if q.is_group_by else None,
                            (got.stats.num_docs_scanned, got.stats.num_total_docs))
    if rank == 0:
        pickle.dump(report, open(os.path.join(out_dir, "report.pkl"), "wb"))
    for d in devs:
        d.destroy()
    dom.release()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("pack_shift,native", [(0, False), (12, False), (0, True), (12, True)])
def test_two_gpus_domain_and_table_reduce_equal_oracle_merge(tmp_path, pack_shift, native):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pinot_b200 import sql
    from gpu_util import assert_tables_equal
    from reduce_util import combine
    port = 33000 + (os.getpid() % 1500) + pack_shift + int(native)
    mp.spawn(_worker, args=(2, port, str(tmp_path), pack_shift, native), nprocs=2, join=True)
    report = pickle.load(open(tmp_path / "report.pkl", "rb"))
    for text, (got, parts, carrier, stats) in report.items():
        q = sql.parse(text)
        assert_tables_equal(q, got, combine([a.function for a in q.aggregations], parts), f"2 GPUs: {text}")
        if native and q.is_group_by and "DISTINCTCOUNT" not in text:
            # the in-library combine also sums the execution statistics over the ranks (the broker's reduce of the metadata)
            assert stats[1] == 2 * (30_000 + 50_001) + 8193 + 8194, stats
        if text.endswith("WHERE v > -400 GROUP BY k, j"):
            # default: 46-bit sum field, safe -> counts carried.  A 12-bit field (values span 2000) cannot be proven safe
            # for the reduce: BOTH ranks agree through the flag all-reduce and rerun without the carrier.
            assert carrier is (pack_shift == 0), (pack_shift, carrier)
