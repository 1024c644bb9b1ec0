#!/bin/bash
# 8 GPUs of one box: in-library combine (one all-reduce carries tables + verdict + statistics) vs the torch-driven reduce, back to back
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
for MODE in native torch; do
  echo "== bench.py --gpus 8 ($MODE)"
  if [ $MODE = torch ]; then export PB200_TORCH_REDUCE=1; else unset PB200_TORCH_REDUCE; fi
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r2_bench_n8_$MODE.json 2> gpurun_out/r2_bench_n8_$MODE.err
  python - <<PY
import json
for line in open("gpurun_out/r2_bench_n8_$MODE.json"):
    if line.startswith("{"):
        j = json.loads(line)
        print({k: j[k] for k in ("n_gpus", "value", "ms_per_step")}, "kernel", j["roofline"]["kernel_ms"], "combine", j.get("cross_gpu_combine"), "e2e", j["e2e"]["value"], "result", j["result"])
PY
  tail -2 gpurun_out/r2_bench_n8_$MODE.err | cut -c1-300
done
