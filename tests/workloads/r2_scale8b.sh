#!/bin/bash
# 8 GPUs of one box: the in-library cross-GPU combine at N = 8 and 4, C4 at N = 8
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for N in 8 4; do
  echo "== bench.py --gpus $N"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2_bench_n${N}_native.json 2> gpurun_out/r2_bench_n${N}_native.err
  python - <<PY
import json
for line in open("gpurun_out/r2_bench_n${N}_native.json"):
    if line.startswith("{"):
        j = json.loads(line)
        print({k: j[k] for k in ("n_gpus", "value", "ms_per_step")}, "kernel", j["roofline"]["kernel_ms"], "combine", j.get("cross_gpu_combine"), "e2e", j["e2e"]["value"], "c2", j["c2"]["ms_per_step"], "result", j["result"])
PY
  tail -2 gpurun_out/r2_bench_n${N}_native.err | cut -c1-300
done
echo "== C4 on 8 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tests/workloads/run_c4.py --segments-per-gpu 8 --rows 50000000 --steps 50 --check > gpurun_out/r2_c4_n8_native.json 2> gpurun_out/r2_c4_n8_native.err
tail -1 gpurun_out/r2_c4_n8_native.json | cut -c1-400; tail -2 gpurun_out/r2_c4_n8_native.err | cut -c1-300
