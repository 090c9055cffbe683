#!/bin/bash
# Round 2, session 3 (2 GPUs): peer-memory transport of the sharded load; data test + bench, both transports.
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/pytest_multi.log 2>&1; tail -8 gpurun_out/pytest_multi.log
for tr in peer; do
RC_SHARD_TRANSPORT=$tr timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/bench_2gpu_$tr.json 2> gpurun_out/bench_2gpu_$tr.err
tail -3 gpurun_out/bench_2gpu_$tr.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_2gpu_$tr.json"))
    print("$tr 2gpu", d["ms_per_step"], d["value"], d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"))
    print(d["run"].get("decomposition"))
except Exception as e: print("failed", e)
PY
done
