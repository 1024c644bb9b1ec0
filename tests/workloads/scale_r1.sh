#!/bin/bash
# multi-GPU evidence: bench.py and the C4 driver under torchrun on N GPUs of one box (N = $1)
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 1500 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/workloads/run_c4.py --segments-per-gpu 8 --rows 50000000 --steps 50 --check > gpurun_out/c4_n$N.json 2> gpurun_out/c4_n$N.err; tail -1 gpurun_out/c4_n$N.json; tail -3 gpurun_out/c4_n$N.err
