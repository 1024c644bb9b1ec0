#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_segment_dir.py tests/test_gpu_startree.py tests/test_gpu_raw.py tests/test_gpu_concurrency.py -m gpu -x -q 2>&1 | tail -8
echo "== C5 card 1M"; timeout 900 python tests/workloads/run_c5.py --rows 100000000 --high-card 1000000 > gpurun_out/r2_c5d.json 2> gpurun_out/r2_c5d.err; cat gpurun_out/r2_c5d.json; tail -3 gpurun_out/r2_c5d.err
