#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== whole gpu suite"; timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "== bench"; timeout 300 python bench.py 2>gpurun_out/r2_bench_final.err | tee gpurun_out/r2_bench_final.json | cut -c1-400; tail -3 gpurun_out/r2_bench_final.err
