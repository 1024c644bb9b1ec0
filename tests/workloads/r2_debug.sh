#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== memcheck smoke"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_memcheck_smoke.log 2>&1; grep -m 40 -E "Invalid|at 0x|by thread|Address|ERROR SUMMARY|in .*\.cu|scan_kernel|kernel" gpurun_out/r2_memcheck_smoke.log | head -60
echo "== pytest parity (first failures)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15
