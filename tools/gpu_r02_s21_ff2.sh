#!/bin/bash
# Packed-pair pilot filter: WBFM parity tests, then cfg3-wbfm / cfg4 bench lines.
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q -k "golden or config4 or config1 or band_plan or config3_literal" > gpurun_out/pytest_ff2.log 2>&1; tail -4 gpurun_out/pytest_ff2.log
for wl in cfg3-wbfm cfg4; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/ff2_$wl.json 2> gpurun_out/ff2_$wl.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ff2_$wl.json"))
    print("FF2 $wl ms/step", round(d["ms_per_step"], 4), "Msps", round(d["value"]), "filtfilt", d["kernels"]["wbfm.pilot_filtfilt"])
except Exception as e: print("FF2 $wl failed", e)
PY
done
