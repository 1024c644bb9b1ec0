#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
echo "== 2-GPU NCCL parity test"; timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -12
