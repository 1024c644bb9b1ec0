#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
echo "== step profile"; timeout 600 python tests/workloads/r2_step_profile.py 2>gpurun_out/r2_step_profile.err | tee gpurun_out/r2_step_profile.json; tail -3 gpurun_out/r2_step_profile.err
echo "== C5 card 1M"; timeout 900 python tests/workloads/run_c5.py --rows 100000000 --high-card 1000000 > gpurun_out/r2_c5b.json 2> gpurun_out/r2_c5b.err; cat gpurun_out/r2_c5b.json; tail -3 gpurun_out/r2_c5b.err
