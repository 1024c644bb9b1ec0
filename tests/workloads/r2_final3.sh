#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== whole gpu suite"; timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
