#!/bin/bash
# tuning sweep on one B200: C2 scan-kernel variants selected by the PB200_* knobs (see pb200_api.cu)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
b() { echo "== $*"; env "$@" python bench.py --quick --steps 100 2>/dev/null | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('ms/step',round(j['ms_per_step'],3),'kernel_ms',round(j['kernel_ms'],4),'frac',round(j['frac'],3),'count',j['count'])"; }
b PB200_X=0
b PB200_W=8 PB200_NO_DEFER=1
b PB200_W=8
b PB200_W=8 PB200_NO_DEFER=1 PB200_STAGES=2
b PB200_W=8 PB200_NO_DEFER=1 PB200_SPARSE_MAX=0
b PB200_SPARSE_MAX=0
