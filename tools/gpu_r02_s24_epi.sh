#!/bin/bash
# One-kernel audio epilogue: parity (golden, config 2, config 4, graph step, fuzz), then cfg2 / cfg4 lines.
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q -k "golden or config2 or config4 or graph_step or band_plan or pipeline" > gpurun_out/pytest_epi.log 2>&1; tail -4 gpurun_out/pytest_epi.log
for wl in cfg2 cfg4; do
  timeout 200 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/epi_$wl.json 2> gpurun_out/epi_$wl.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/epi_$wl.json"))
    print("EPI $wl ms/step", round(d["ms_per_step"], 4), "Msps", round(d["value"]), {k: v["avg_ms"] for k, v in d["kernels"].items() if "deemph" in k or "mean" in k})
except Exception as e: print("EPI $wl failed", e)
PY
done
