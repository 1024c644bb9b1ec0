#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
show() { python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(j['query'][:80].ljust(80), j['variant'], 'step', round(j['ms_per_step'],3), 'kernel', round(j['scan_kernel_ms'],3), 'frac', round(j['frac_of_peak'],3), j['same_as_first_variant'], j['checked'])
"; }
{
timeout 600 python tests/workloads/run_c3.py --mode range --steps 10 --check-rows 1000000 --variants "s1:table_stride=1 auto: s4:table_stride=4 s16:table_stride=16 s32:table_stride=32 nopack_s1:pack_count=0,table_stride=1 nopack_auto:pack_count=0" 2>&1 | show
timeout 600 python tests/workloads/run_c3.py --mode range2 --steps 5 --variants "s1:table_stride=1 auto:" 2>&1 | show
for v in "PB200_TABLE_STRIDE=1" "PB200_X=0"; do echo "== C4 $v"; env $v timeout 300 python tests/workloads/run_c4.py --check 2>&1 | tail -1 | cut -c1-300; done
for v in "PB200_TABLE_STRIDE=1" "PB200_X=0" "PB200_TABLE_STRIDE=32"; do echo "== bench --quick $v"; env $v timeout 600 python bench.py --quick --steps 50 2>&1 | tail -1 | cut -c1-330; done
} | tee gpurun_out/r2_stride.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_datatable.py -m gpu -q -x 2>&1 | tail -4
