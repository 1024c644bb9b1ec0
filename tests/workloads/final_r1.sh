#!/bin/bash
# end-of-round evidence on one B200: GPU tests, smoke, bench (both arms), ncu launch list + full capture of the C2 kernel
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_n1.json 2> gpurun_out/bench_reference_n1.err; tail -c 600 gpurun_out/bench_reference_n1.json
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/launches_r1.log 2>&1
bash tests/workloads/profile_r1.sh prof_r1_c2_final
