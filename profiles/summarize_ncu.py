#!/usr/bin/env python
"""Summarises one `ncu --set full --import-source on` capture of pb200::scan_kernel into the text kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/<report>.ncu-rep [warp_tiles] > profiles/<name>.txt

`warp_tiles` = number of 1024-row warp slices the launch processed (rows / 1024; 781250 for the 800 M-row C2 / C3 launches):
per-region instruction counts are printed per warp slice.  Reads the report with `ncu -i ... --page raw|source --csv`.
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
WT = float(sys.argv[2]) if len(sys.argv) > 2 else 781250
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "lts__t_sectors.sum", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# {rep}")
print("## raw metrics (one launch)")
for i, h in enumerate(hdr):
    if h in ("Kernel Name",):
        print(h, vals[i])
    if h in want or ("average_warps_issue_stalled" in h and "per_issue_active" in h and float(vals[i] or 0) > 0.1):
        print(h, units[i], vals[i])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
isrc, iex, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [(r[isrc].strip(), int(r[iex]), int(r[isamp]), r) for r in rows[2:] if len(r) > iex]
tot = sum(d[1] for d in data)
ts = sum(d[2] for d in data)
print(f"\n## SASS regions executed at least 0.25x per 1024-row warp slice, by stall samples")
print(f"total warp-instructions {tot} = {tot / WT:.1f} per warp slice; {ts} samples; {len(data)} SASS instructions in the kernel")
regions, cur = [], None
for idx, d in enumerate(data):
    if d[1] >= 0.25 * WT:
        if cur is None:
            cur = [idx, idx, 0, 0]
        cur[1] = idx
        cur[2] += d[1]
        cur[3] += d[2]
    elif cur is not None:
        regions.append(cur)
        cur = None
if cur:
    regions.append(cur)
cold = ts - sum(r[3] for r in regions)
print(f"samples outside these regions (code run for a subset of the slices / rows): {cold} ({100 * cold / max(ts, 1):.1f} %)")
regions.sort(key=lambda r: -r[3])
for r in regions[:16]:
    ops, st = {}, {}
    for d in data[r[0]:r[1] + 1]:
        tok = d[0].split()
        op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
        for c in stall_cols:
            v = int(d[3][c] or 0)
            if v:
                st[hdr[c]] = st.get(hdr[c], 0) + v
    top = ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
    tst = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(f"SASS #{r[0]}-{r[1]}: {r[2] / WT:.0f} instr/slice, {100 * r[3] / ts:.1f} % of samples | ops: {top} | stalls: {tst} | first: {data[r[0]][0][:40]}")
