#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
echo "== 2-GPU NCCL parity test"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6
echo "== combine profile"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/workloads/r2_combine_profile.py 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/r2_combine_profile_n2.json
