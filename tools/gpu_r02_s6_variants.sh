#!/bin/bash
# Sharded-load transport variants at N = <gpus>: where should the exchanges live?
G=${1:-2}
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
run() {
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G --steps 10 --warmup 3 --no-extras --no-e2e > gpurun_out/var_${G}gpu_$tag.json 2> gpurun_out/var_${G}gpu_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/var_${G}gpu_$tag.json"))
    print("VARIANT $tag ${G}gpu ms/step", round(d["ms_per_step"], 4), "Msps", round(d["value"]), d["run"].get("decomposition"))
except Exception as e: print("VARIANT $tag failed", e)
PY
}
run copies RC_SHARD_FUSED=0
run fused_combine_cta1 RC_SHARD_FUSED=combine RC_SCATTER_CTAS=1
run fused_combine_cta2 RC_SHARD_FUSED=combine RC_SCATTER_CTAS=2
run fused_combine_cta4 RC_SHARD_FUSED=combine RC_SCATTER_CTAS=4
run fused_both_cta2 RC_SHARD_FUSED=1 RC_SCATTER_CTAS=2
run fused_fft RC_SHARD_FUSED=fft
