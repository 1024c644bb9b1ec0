#!/bin/bash
# one-box regression + workload round: GPU tests, C2 sweep head, C3 / C4 / C5 drivers
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tests/workloads/sweep_r1.sh 2>&1 | head -2
python tests/workloads/run_c3.py --mode bitmap --steps 10 --check-rows 2000000
python tests/workloads/run_c3.py --mode range --steps 10 --check-rows 2000000
PB200_NO_SMEM_GROUPS=1 python tests/workloads/run_c3.py --mode range --steps 10
python tests/workloads/run_c3.py --mode range2 --steps 5
python tests/workloads/run_c4.py --check 2>&1 | tail -1
if [ "$1" = "c5" ]; then timeout 400 python tests/workloads/run_c5.py --rows 100000000 2>&1 | tail -1; fi
