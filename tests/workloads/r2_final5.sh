#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_cli.py tests/test_gpu_filtered.py -m gpu -q -x 2>&1 | tail -15
