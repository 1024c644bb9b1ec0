#!/bin/bash
# round 2, call 1: GPU parity with the software-pipelined group-by, C3/range + C4 variants, one ncu capture
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r2_pytest1.log
tail -3 gpurun_out/r2_pytest1.log
V="base:gb_defer=0 defer: defer_c1:ctas_per_sm=1 defer_w8:warps=8 base_w8:gb_defer=0,warps=8 defer_s2:stages=2 smem10k:smem_groups_max=10000,ctas_per_sm=1"
timeout 600 python tests/workloads/run_c3.py --mode range --steps 10 --check-rows 1000000 --variants "$V" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(j['variant'], round(j['ms_per_step'],3), round(j['scan_kernel_ms'],3), round(j['frac_of_peak'],3), j['same_as_first_variant'], j['checked'])
" | tee gpurun_out/r2_c3range_variants.log
timeout 600 python tests/workloads/run_c3.py --mode range2 --steps 5 --variants "base:gb_defer=0 defer:" 2>&1 | cut -c1-400 | tee gpurun_out/r2_c3range2.log
timeout 600 python tests/workloads/run_c3.py --mode bitmap --steps 10 --variants "base:gb_defer=0 defer:" 2>&1 | cut -c1-400 | tee gpurun_out/r2_c3bitmap.log
for v in "PB200_NO_GB_DEFER=1" "PB200_X=0"; do echo "== C4 $v"; env $v timeout 300 python tests/workloads/run_c4.py --check 2>&1 | tail -1 | cut -c1-300; done | tee gpurun_out/r2_c4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_c3range_a python tests/workloads/run_c3.py --mode range --steps 2 --warmup 1 > gpurun_out/prof_r2_c3range_a.log 2>&1; tail -1 gpurun_out/prof_r2_c3range_a.log | cut -c1-200
