#!/bin/bash
# Round 2, multi-GPU session: tools/gpu_r02_s5_multi.sh <gpus>
# data test for every world size the box allows, then the driver's own command line at N = <gpus>
# (default mode: one stream, sharded Tuner.load; extra keys bcast / replicas / cfg5 at 8 GPUs).
G=${1:-2}
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi topo -m > gpurun_out/topo_${G}gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -k "$G" > gpurun_out/pytest_multi_${G}gpu.log 2>&1; tail -8 gpurun_out/pytest_multi_${G}gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $G --steps 20 --warmup 5 > gpurun_out/bench_${G}gpu.json 2> gpurun_out/bench_${G}gpu.err
tail -3 gpurun_out/bench_${G}gpu.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${G}gpu.json"))
    print("${G}gpu", d["ms_per_step"], d["value"], d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"))
    print(d["run"].get("decomposition"))
    for k in ("bcast", "replicas", "cfg5"): print(k, d.get(k))
except Exception as e: print("failed", e)
PY
