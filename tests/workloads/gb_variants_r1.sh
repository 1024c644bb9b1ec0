#!/bin/bash
# group-by kernel variants (tuning knobs of pb200_api.cu) on C3/range (10 000 groups), range3 (100 groups) and C4
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
r() { mode=$1; shift; echo "== $mode $*"; env "$@" python tests/workloads/run_c3.py --mode $mode --steps 10 --check-rows 1000000 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_step'],3), round(j['scan_kernel_ms'],3), j['checked'])"; }
r range PB200_X=0
r range PB200_CTAS=1
r range PB200_W=8
r range3 PB200_X=0
r range3 PB200_CTAS=1
r bitmap PB200_X=0
r range2 PB200_X=0
for v in "PB200_X=0" "PB200_W=8"; do echo "== C4 $v"; env $v python tests/workloads/run_c4.py --check 2>&1 | tail -1 | cut -c1-250; done
if [ "$1" = "prof" ]; then ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -f -o gpurun_out/prof_r1_c3range python tests/workloads/run_c3.py --mode range --steps 2 --warmup 1 > gpurun_out/prof_r1_c3range.log 2>&1; tail -1 gpurun_out/prof_r1_c3range.log | cut -c1-150; fi
