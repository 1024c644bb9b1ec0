#!/bin/bash
# C2 scan kernel across selectivities (bench.py --quick --selectivity S)
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
b() { sel=$1; shift; echo "== sel $sel $*"; env "$@" python bench.py --quick --steps 60 --selectivity $sel 2>/dev/null | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('ms/step',round(j['ms_per_step'],3),'kernel_ms',round(j['kernel_ms'],4),'frac',round(j['frac'],3))"; }
for sel in ${SELS:-0.0001 0.001 0.01 0.05 0.1 0.25 0.5 1.0}; do
  b $sel PB200_X=0
done
