#!/bin/bash
# Pipeline lanes of the sharded load: data test at world <gpus>, then bench with 1 and 2 lanes.
G=${1:-2}
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -k "$G" > gpurun_out/pytest_multi_${G}gpu.log 2>&1; tail -4 gpurun_out/pytest_multi_${G}gpu.log
for L in 2 1; do
RC_SHARD_LANES=$L timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G --steps 20 --warmup 5 --no-extras > gpurun_out/lanes${L}_${G}gpu.json 2> gpurun_out/lanes${L}_${G}gpu.err
tail -2 gpurun_out/lanes${L}_${G}gpu.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/lanes${L}_${G}gpu.json"))
    print("LANES $L ${G}gpu ms/step", round(d["ms_per_step"], 4), "Msps", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["run"].get("decomposition"))
    print(d["roofline"])
except Exception as e: print("LANES $L failed", e)
PY
done
