#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
echo "== raw column tests"; timeout 900 python -m pytest tests/test_gpu_raw.py -m gpu -x -q 2>&1 | tail -15
echo "== whole gpu suite"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
echo "== step profile"; timeout 600 python tests/workloads/r2_step_profile.py 2>gpurun_out/r2_step_profile.err | tee gpurun_out/r2_step_profile2.json; tail -3 gpurun_out/r2_step_profile.err
echo "== C4 N=1"; timeout 600 python tests/workloads/run_c4.py 2>&1 | grep '^{' | tee gpurun_out/r2_c4_n1.json
echo "== C5 card 1M"; timeout 900 python tests/workloads/run_c5.py --rows 100000000 --high-card 1000000 > gpurun_out/r2_c5c.json 2> gpurun_out/r2_c5c.err; cat gpurun_out/r2_c5c.json; tail -3 gpurun_out/r2_c5c.err
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 2>gpurun_out/r2_bench_b.err | tee gpurun_out/r2_bench_b.json | cut -c1-600; tail -3 gpurun_out/r2_bench_b.err
