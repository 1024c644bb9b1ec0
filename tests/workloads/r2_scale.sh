#!/bin/bash
# multi-GPU evidence on ONE box: gpurun --gpus N -- bash tests/workloads/r2_scale.sh N
N=${1:-2}
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$N" = "2" ]; then
  echo "== 2-GPU NCCL parity test"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6
fi
echo "== bench.py --gpus $N"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 2500 gpurun_out/r2_bench_n$N.json; tail -3 gpurun_out/r2_bench_n$N.err
echo "== C4 on $N GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tests/workloads/run_c4.py --segments-per-gpu 8 --rows 50000000 --steps 50 --check > gpurun_out/r2_c4_n$N.json 2> gpurun_out/r2_c4_n$N.err
tail -1 gpurun_out/r2_c4_n$N.json | cut -c1-400; tail -2 gpurun_out/r2_c4_n$N.err
if [ "$N" = "2" ]; then
  echo "== ncu full: aggregation-only kernel (c2), one GPU"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_bench_c2 python tests/workloads/run_c2.py --steps 2 --warmup 2 > gpurun_out/prof_r2_bench_c2.log 2>&1; tail -1 gpurun_out/prof_r2_bench_c2.log | cut -c1-200
  echo "== segment cache / datatable tests"; timeout 600 python -m pytest tests/test_gpu_segment_dir.py tests/test_gpu_datatable.py -m gpu -q 2>&1 | tail -3
fi
