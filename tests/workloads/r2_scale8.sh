#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== 2-GPU NCCL parity test (verdict slot)"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4
for N in 8 4 2; do
  echo "== bench.py --gpus $N"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
  python - <<PY
import json
for line in open("gpurun_out/r2_bench_n$N.json"):
    if line.startswith("{"):
        j = json.loads(line)
        print({k: j[k] for k in ("n_gpus", "value", "ms_per_step")}, "kernel", j["roofline"]["kernel_ms"], "carrier", j["count_carrier"], "e2e", j["e2e"]["value"], "c2", j["c2"]["ms_per_step"], "result", j["result"])
PY
  tail -2 gpurun_out/r2_bench_n$N.err | cut -c1-300
done
echo "== C4 on 8 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tests/workloads/run_c4.py --segments-per-gpu 8 --rows 50000000 --steps 50 --check > gpurun_out/r2_c4_n8.json 2> gpurun_out/r2_c4_n8.err
tail -1 gpurun_out/r2_c4_n8.json | cut -c1-400; tail -2 gpurun_out/r2_c4_n8.err | cut -c1-300
echo "== combine profile 8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 tests/workloads/r2_combine_profile.py 2>&1 | grep -E "^\{" | tee gpurun_out/r2_combine_profile_n8.json
