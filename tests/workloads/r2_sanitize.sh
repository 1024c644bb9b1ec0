#!/bin/bash
# compute-sanitizer over the round-2 kernels (count carrier, extraction, re-encode, verify) + C5 at the SURVEY's cardinality
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
K='count_carrier or long_sum or (random_segments and 4096) or (hash_group_table_parity and 33) or raw_forward or golden_inner_segment_group_by'
echo "== memcheck"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_domain.py tests/test_gpu_datatable.py -m gpu -x -q -k "$K or domain or datatable" > gpurun_out/r2_memcheck.log 2>&1; echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_memcheck.log | tail -4
echo "== racecheck"; timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_domain.py -m gpu -x -q -k "count_carrier or (random_segments and 4096) or (hash_group_table_parity and 33) or domain_merge" > gpurun_out/r2_racecheck.log 2>&1; echo "exit $?"; grep -E "RACECHECK SUMMARY|hazards|passed|failed" gpurun_out/r2_racecheck.log | tail -4
echo "== C5, d_hi cardinality 1 000 000"; timeout 1500 python tests/workloads/run_c5.py --rows 100000000 --high-card 1000000 > gpurun_out/r2_c5.json 2> gpurun_out/r2_c5.err; tail -c 2500 gpurun_out/r2_c5.json; tail -3 gpurun_out/r2_c5.err
