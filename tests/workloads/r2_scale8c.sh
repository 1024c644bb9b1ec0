#!/bin/bash
# 8 GPUs of one box: in-library combine (one all-reduce carries tables + verdict + statistics) vs the torch-driven reduce, back to back
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
echo "== 2-GPU NCCL parity test"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -4
for MODE in native torch; do
  echo "== bench.py --gpus 8 ($MODE)"
  if [ $MODE = torch ]; then export PB200_TORCH_REDUCE=1; else unset PB200_TORCH_REDUCE; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r2_bench_n8_$MODE.json 2> gpurun_out/r2_bench_n8_$MODE.err
  python - <<PY
import json
for line in open("gpurun_out/r2_bench_n8_$MODE.json"):
    if line.startswith("{"):
        j = json.loads(line)
        print({k: j[k] for k in ("n_gpus", "value", "ms_per_step")}, "kernel", j["roofline"]["kernel_ms"], "combine", j.get("cross_gpu_combine"), "e2e", j["e2e"]["value"], "result", j["result"])
PY
  tail -2 gpurun_out/r2_bench_n8_$MODE.err | cut -c1-300
done
unset PB200_TORCH_REDUCE
echo "== C4 on 8 GPUs"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tests/workloads/run_c4.py --segments-per-gpu 8 --rows 50000000 --steps 50 --check > gpurun_out/r2_c4_n8_native.json 2> gpurun_out/r2_c4_n8_native.err
tail -1 gpurun_out/r2_c4_n8_native.json | cut -c1-400
