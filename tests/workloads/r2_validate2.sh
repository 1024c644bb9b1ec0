#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
echo "== filtered aggregation tests (1 GPU)"; timeout 300 python -m pytest tests/test_gpu_filtered.py -m gpu -x -q 2>&1 | tail -12
echo "== 2-GPU NCCL parity test"; timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -6
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== bench N=2, in-library combine"; timeout 240 bash -c "$(declare -f run); run 29551 bench.py --gpus 2 --steps 50 --warmup 5" 2>gpurun_out/r2_native_n2.err | grep '^{' | tee gpurun_out/r2_bench_n2_native.json | cut -c1-300
tail -3 gpurun_out/r2_native_n2.err
echo "== C4 N=2"; timeout 200 bash -c "$(declare -f run); run 29553 tests/workloads/run_c4.py" 2>&1 | grep '^{' | tee gpurun_out/r2_c4_n2_native.json
