#!/bin/bash
# Round 2, session 2 (2 GPUs): multi-GPU data test, then the driver's own N=2 command line.
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/pytest_multi.log 2>&1; tail -15 gpurun_out/pytest_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -5 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_2gpu.json"))
    print("2gpu", d["ms_per_step"], d["value"], d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"))
    for k, v in d["kernels"].items(): print("   ", k, v["avg_ms"], v["frac_of_hbm_peak"])
    for k in ("bcast", "replicas"): print(k, d.get(k))
except Exception as e: print("failed", e)
PY
