#!/bin/bash
# round 2, call 2: per-thread sparse group-by path; domain tests; new extraction
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest2.log 2>&1; tail -5 gpurun_out/r2_pytest2.log
show() { python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(j['workload'], j['variant'], 'step', round(j['ms_per_step'],3), 'kernel', round(j['scan_kernel_ms'],3), 'frac', round(j['frac_of_peak'],3), j['same_as_first_variant'], j['checked'])
"; }
V="queue:sparse_max_gb=0 sparse12: sparse8:sparse_max_gb=8 sparse32:sparse_max_gb=32 sparse12_w8:warps=8 sparse12_c1:ctas_per_sm=1 sparse12_s2:stages=2"
timeout 600 python tests/workloads/run_c3.py --mode range --steps 10 --check-rows 1000000 --variants "$V" 2>&1 | show | tee gpurun_out/r2_c3range_variants2.log
timeout 600 python tests/workloads/run_c3.py --mode range2 --steps 5 --variants "queue:sparse_max_gb=0 sparse12:" 2>&1 | show | tee gpurun_out/r2_c3range2_b.log
timeout 600 python tests/workloads/run_c3.py --mode bitmap --steps 10 --variants "queue:sparse_max_gb=0 sparse12:" 2>&1 | show | tee gpurun_out/r2_c3bitmap_b.log
for v in "PB200_SPARSE_MAX_GB=0" "PB200_X=0" "PB200_W=8"; do echo "== C4 $v"; env $v timeout 300 python tests/workloads/run_c4.py --check 2>&1 | tail -1 | cut -c1-300; done | tee gpurun_out/r2_c4_b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_c3range_b python tests/workloads/run_c3.py --mode range --steps 2 --warmup 1 > gpurun_out/prof_r2_c3range_b.log 2>&1; tail -1 gpurun_out/prof_r2_c3range_b.log | cut -c1-200
