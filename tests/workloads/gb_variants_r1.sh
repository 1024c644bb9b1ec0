#!/bin/bash
# group-by kernel variants on C3/range and C4 (tuning knobs of pb200_api.cu)
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in "PB200_X=0" "PB200_NO_SMEM_GROUPS=1" "PB200_CTAS=2"; do echo "== $v"; env $v python tests/workloads/run_c3.py --mode range --steps 10 --check-rows 1000000 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['scan_kernel_ms'], j['checked'])"; done
python tests/workloads/run_c3.py --mode bitmap --steps 10 --check-rows 1000000 | cut -c1-400
for v in "PB200_X=0" "PB200_CTAS=2"; do echo "== C4 $v"; env $v python tests/workloads/run_c4.py --check 2>&1 | tail -1 | cut -c1-250; done
if [ "$1" = "prof" ]; then ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -f -o gpurun_out/prof_r1_c3range python tests/workloads/run_c3.py --mode range --steps 2 --warmup 1 > gpurun_out/prof_r1_c3range.log 2>&1; tail -1 gpurun_out/prof_r1_c3range.log | cut -c1-150; fi
