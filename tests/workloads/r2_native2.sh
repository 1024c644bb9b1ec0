#!/bin/bash
# 2 GPUs: the in-library NCCL combine (pb200_result_combine) against the torch-driven reduce
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
echo "== 2-GPU NCCL parity test (torch-driven and in-library combine)"; timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -8
echo "== star-tree tests"; timeout 600 python -m pytest tests/test_gpu_startree.py -m gpu -q 2>&1 | tail -3
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== bench N=2, in-library combine"; timeout 900 bash -c "$(declare -f run); run 29551 bench.py --gpus 2 --steps 20 --warmup 5" 2>gpurun_out/r2_native_n2.err | grep '^{' | tee gpurun_out/r2_bench_n2_native.json | cut -c1-300
echo "== bench N=2, torch-driven reduce"; PB200_TORCH_REDUCE=1 timeout 900 bash -c "$(declare -f run); run 29552 bench.py --gpus 2 --steps 20 --warmup 5" 2>gpurun_out/r2_torch_n2.err | grep '^{' | tee gpurun_out/r2_bench_n2_torch.json | cut -c1-300
echo "== C4 N=2 native"; timeout 600 bash -c "$(declare -f run); run 29553 tests/workloads/run_c4.py" 2>&1 | grep '^{' | tee gpurun_out/r2_c4_n2_native.json
echo "== C4 N=2 torch"; PB200_TORCH_REDUCE=1 timeout 600 bash -c "$(declare -f run); run 29554 tests/workloads/run_c4.py" 2>&1 | grep '^{' | tee gpurun_out/r2_c4_n2_torch.json
tail -5 gpurun_out/r2_native_n2.err
