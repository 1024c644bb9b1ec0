#!/bin/bash
# quick end-to-end validation of the current tree on one B200: GPU tests, smoke, full bench line
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json; j=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print('value',j['value'],'ms',j['ms_per_step'],'frac',j['roofline']['frac'],'e2e',j['e2e']['value'],'cpu',j['cpu_baseline']['value'])"
