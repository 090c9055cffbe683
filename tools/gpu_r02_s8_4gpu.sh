#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -k 4 > gpurun_out/pytest_multi_4gpu.log 2>&1; tail -4 gpurun_out/pytest_multi_4gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
tail -3 gpurun_out/bench_4gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_4gpu.json"))
print("4gpu", d["ms_per_step"], d["value"], d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"))
print(d["run"].get("decomposition"))
for k in ("bcast", "replicas"): print(k, d.get(k))
PY
