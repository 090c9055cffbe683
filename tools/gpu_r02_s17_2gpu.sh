#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 examples/multi_gpu_sharded.py 3 > gpurun_out/example_sharded_2gpu.log 2>&1; tail -3 gpurun_out/example_sharded_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_2gpu.json"))
print("2gpu", d["ms_per_step"], d["value"], d.get("e2e", {}).get("value"), d["run"].get("decomposition"))
print(d["roofline"])
PY
