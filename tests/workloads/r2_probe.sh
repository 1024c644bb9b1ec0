#!/bin/bash
# what bounds the group-by kernel: marginal cost of each per-survivor operation (C3 table, 8 x 100 M rows)
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest3.log 2>&1; tail -3 gpurun_out/r2_pytest3.log
show() { python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(j['query'][:95].ljust(95), j['variant'], 'step', round(j['ms_per_step'],3), 'kernel', round(j['scan_kernel_ms'],3), 'GB/s', round(j['achieved_gbs']), 'matched', j['matched'])
"; }
R() { timeout 300 python tests/workloads/run_c3.py --mode range --steps 8 --query "$1" --bits $2 --variants "${3:-q:}" 2>&1 | show; }
{
R "SELECT COUNT(*) FROM t WHERE f BETWEEN 3002 AND 5999" 14
R "SELECT COUNT(*) FROM t WHERE f BETWEEN 3002 AND 5999 GROUP BY g" 28
R "SELECT SUM(m) FROM t WHERE f BETWEEN 3002 AND 5999 GROUP BY g" 45
R "SELECT MAX(m) FROM t WHERE f BETWEEN 3002 AND 5999 GROUP BY g" 45
R "SELECT SUM(m), COUNT(*) FROM t WHERE f BETWEEN 3002 AND 5999 GROUP BY g" 45 "q: s12:sparse_max_gb=12"
R "SELECT SUM(m), COUNT(*), MAX(f) FROM t WHERE f BETWEEN 3002 AND 5999 GROUP BY g" 45
R "SELECT SUM(m), COUNT(*) FROM t WHERE f BETWEEN 3002 AND 3299 GROUP BY g" 45
R "SELECT SUM(m), COUNT(*) FROM t WHERE f BETWEEN 3002 AND 4499 GROUP BY g" 45
R "SELECT SUM(m), COUNT(*) FROM t WHERE f BETWEEN 3002 AND 8999 GROUP BY g" 45
R "SELECT SUM(m), COUNT(*) FROM t WHERE f BETWEEN 3002 AND 17999 GROUP BY g" 45
R "SELECT SUM(m), COUNT(*) FROM t WHERE f BETWEEN 3002 AND 5999 GROUP BY g2" 38
R "SELECT SUM(m), COUNT(*) FROM t WHERE f BETWEEN 3002 AND 5999 GROUP BY d3" 37
} | tee gpurun_out/r2_probe.log
