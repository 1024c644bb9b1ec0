#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
echo "== filtered aggregation + raw tests"; timeout 120 python -m pytest tests/test_gpu_filtered.py tests/test_gpu_raw.py -m gpu -q -x 2>&1 | tail -4
echo "== launch list of bench.py --quick"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_raw.csv python bench.py --quick --steps 3 --warmup 3 > gpurun_out/r2_launches.log 2>&1
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
with open("gpurun_out/r2_launch_list_final.csv", "w") as f:
    f.write("kernel,launches,total_ms,avg_ms\n")
    for k, (n, ms) in agg.items():
        f.write(f"\"{k}\",{n},{ms:.4f},{ms / n:.4f}\n")
print(open("gpurun_out/r2_launch_list_final.csv").read())
PY
