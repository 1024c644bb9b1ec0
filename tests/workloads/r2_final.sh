#!/bin/bash
# round 2 evidence run on ONE GPU: tests, the bench line (both arms), launch list, ncu captures of both kernels
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; tail -4 gpurun_out/r2_pytest_final.log
echo "== bench.py (b200 arm)"; timeout 1500 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 3000 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_bench_n1.err
echo "== bench.py --impl reference"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err; cut -c1-600 gpurun_out/r2_bench_reference_n1.json; tail -3 gpurun_out/r2_bench_reference_n1.err
echo "== launch list (bench --quick, 3 timed steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_raw.csv python bench.py --quick --steps 3 --warmup 3 > gpurun_out/r2_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_launches_raw.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][:90]
    try: ns = float(r[-1].replace(",", ""))
    except ValueError: continue
    unit = r[-2]
    ms = ns / 1e6 if unit in ("ns", "nsecond") else ns / 1e3 if unit in ("us", "usecond") else ns
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
with open("gpurun_out/r2_launch_list.csv", "w") as f:
    f.write("kernel,launches,total_ms,avg_ms\n")
    for k, (n, ms) in agg.items():
        f.write(f"\"{k}\",{n},{ms:.4f},{ms / n:.4f}\n")
print(open("gpurun_out/r2_launch_list.csv").read())
PY
echo "== ncu full: group-by kernel of the bench headline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 6 -c 1 -f -o gpurun_out/prof_r2_bench_gb python bench.py --quick --steps 3 --warmup 3 > gpurun_out/prof_r2_bench_gb.log 2>&1; tail -1 gpurun_out/prof_r2_bench_gb.log | cut -c1-200
echo "== ncu full: aggregation-only kernel (c2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -f -o gpurun_out/prof_r2_bench_c2 python tests/workloads/run_c2.py --steps 2 --warmup 2 > gpurun_out/prof_r2_bench_c2.log 2>&1; tail -1 gpurun_out/prof_r2_bench_c2.log | cut -c1-200
