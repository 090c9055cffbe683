#!/bin/bash
G=${1:-2}
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
run() {
  tag=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G --steps 20 --warmup 5 --no-extras --no-e2e > gpurun_out/prio_${G}gpu_$tag.json 2> gpurun_out/prio_${G}gpu_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/prio_${G}gpu_$tag.json"))
    print("VARIANT $tag ${G}gpu ms/step", round(d["ms_per_step"], 4), "Msps", round(d["value"]), d["run"].get("decomposition"))
except Exception as e: print("VARIANT $tag failed", e)
PY
}
run high RC_SHARD_PRIORITY=1
run high_cta4 RC_SHARD_PRIORITY=1 RC_SCATTER_CTAS=4
run high_noturns RC_SHARD_PRIORITY=1 RC_SHARD_TURNS=0
run default RC_SHARD_PRIORITY=0
