#!/bin/bash
# ncu captures of the C2 scan kernel (one launch each); numbers printed under ncu are NOT bench values
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
prof() { name=$1; shift; env "$@" ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -f -o gpurun_out/$name python bench.py --quick --steps 3 --warmup 3 > gpurun_out/$name.log 2>&1; tail -2 gpurun_out/$name.log | cut -c1-200; }
prof ${1:-prof_r1_c2} PB200_X=0
