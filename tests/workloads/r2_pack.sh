#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest4.log 2>&1; tail -4 gpurun_out/r2_pytest4.log
show() { python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(j['query'][:90].ljust(90), j['variant'], 'step', round(j['ms_per_step'],3), 'kernel', round(j['scan_kernel_ms'],3), 'frac', round(j['frac_of_peak'],3), j['same_as_first_variant'], j['checked'])
"; }
{
timeout 600 python tests/workloads/run_c3.py --mode range --steps 10 --check-rows 1000000 --variants "nopack:pack_count=0 pack: pack_w8:warps=8 pack_s2:stages=2" 2>&1 | show
timeout 600 python tests/workloads/run_c3.py --mode bitmap --steps 10 --variants "nopack:pack_count=0 pack:" 2>&1 | show
timeout 600 python tests/workloads/run_c3.py --mode range2 --steps 5 --variants "nopack:pack_count=0 pack:" 2>&1 | show
for v in "PB200_NO_PACK_COUNT=1" "PB200_X=0"; do echo "== C4 $v"; env $v timeout 300 python tests/workloads/run_c4.py --check 2>&1 | tail -1 | cut -c1-300; done
echo "== bench --quick"; timeout 600 python bench.py --quick --steps 50 2>&1 | tail -1 | cut -c1-400
echo "== bench --quick nopack"; PB200_NO_PACK_COUNT=1 timeout 600 python bench.py --quick --steps 50 2>&1 | tail -1 | cut -c1-400
} | tee gpurun_out/r2_pack.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_c3range_c python tests/workloads/run_c3.py --mode range --steps 2 --warmup 1 > gpurun_out/prof_r2_c3range_c.log 2>&1; tail -1 gpurun_out/prof_r2_c3range_c.log | cut -c1-200
