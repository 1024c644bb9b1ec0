"""-m gpu: the thread-safety contract of include/pinot_b200.h -- one pb200_ctx shared by many host threads (Pinot runs one
operator per worker thread, concurrently for many queries, while segments are loaded and dropped: ServerQueryExecutorV1Impl
+ TableDataManager).  12 threads hammer ONE context with pb200_execute on shared resident segments and with
pb200_segment_register / pb200_segment_release cycles of their own; every result is compared with the CPU oracle."""
import threading

import numpy as np
import pytest

from gpu_util import assert_tables_equal, gpu_table, oracle_table, to_device
from pinot_b200 import sql
from pinot_b200.plan_maker import B200Context, B200PlanMaker

pytestmark = pytest.mark.gpu

QUERIES = [
    "SELECT COUNT(*), SUM(b), MIN(c), MAX(c) FROM t WHERE b > 300",
    "SELECT SUM(c), COUNT(*) FROM t WHERE b BETWEEN 100 AND 600 AND c > 20000",
    "SELECT COUNT(*), SUM(b) FROM t WHERE a = 1 GROUP BY d",                       # inverted index + shared-memory tables
    "SELECT SUM(c), MAX(b), AVG(b) FROM t WHERE b < 800 GROUP BY d, a",             # global tables
    "SELECT DISTINCTCOUNT(b), COUNT(*) FROM t WHERE c > 100000",
    "SELECT COUNT(*), MAX(c) FROM t WHERE b NOT IN (1, 5, 9, 500) GROUP BY a",
]


def _segment(oracle, seed, n):
    rng = np.random.default_rng(seed)
    return oracle.build_segment(f"cc{seed}", {
        "a": rng.integers(0, 7, size=n).astype(np.int32),
        "b": rng.integers(0, 1000, size=n).astype(np.int32),
        "c": rng.integers(0, 70000, size=n).astype(np.int32) * 11,
        "d": rng.integers(0, 200, size=n).astype(np.int32)}, inverted=["a"])


def test_concurrent_execute_and_segment_lifecycle(oracle):
    ctx = B200Context(0)
    pm = B200PlanMaker(ctx)
    qs = [sql.parse(t) for t in QUERIES]
    shared = [_segment(oracle, 900 + i, n) for i, n in enumerate([50_000, 8193, 120_000])]
    shared_dev = [to_device(ctx, s) for s in shared]
    want_shared = [[oracle_table(s, q, oracle.execute(s, q)) for q in qs] for s in shared]
    private = [_segment(oracle, 700 + t, 20_000 + 1111 * t) for t in range(12)]
    want_private = [[oracle_table(s, q, oracle.execute(s, q)) for q in qs] for s in private]
    errors = []
    start = threading.Barrier(12)

    def worker(t):
        try:
            start.wait()
            for it in range(6):
                dev = to_device(ctx, private[t])                     # pb200_segment_register while others execute
                for k in range(len(qs)):
                    qi = (k + t + it) % len(qs)
                    si = (t + k) % len(shared)
                    blk = pm.make_segment_plan_node(shared_dev[si], qs[qi]).run().next_block()
                    assert_tables_equal(qs[qi], gpu_table(shared[si], qs[qi], blk), want_shared[si][qi], f"thread {t} shared {si} q{qi}")
                    blk = pm.make_segment_plan_node(dev, qs[qi]).run().next_block()
                    assert_tables_equal(qs[qi], gpu_table(private[t], qs[qi], blk), want_private[t][qi], f"thread {t} private q{qi}")
                # several segments in one submission, from several threads at once
                blocks = pm.execute_segments([shared_dev[0], dev, shared_dev[2]], qs[it % len(qs)])
                for sd, w, b in zip((shared[0], private[t], shared[2]), (want_shared[0], want_private[t], want_shared[2]), blocks):
                    assert_tables_equal(qs[it % len(qs)], gpu_table(sd, qs[it % len(qs)], b), w[it % len(qs)], f"thread {t} batch")
                dev.destroy()                                        # pb200_segment_release
        except BaseException as e:  # noqa: BLE001 -- reported to the main thread
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(12)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for d in shared_dev:
        d.destroy()
    ctx.close()
    assert not errors, errors[:3]
