#!/bin/bash
# Round 2, session 13 (1 GPU): compute-sanitizer over every kernel family incl. the new ones, the
# full GPU suite, the small configs after the tuned splits.
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 300 python tools/sanitize_run.py > gpurun_out/sanitize_plain.log 2>&1; tail -6 gpurun_out/sanitize_plain.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitize_$tool.log 2>&1
  tail -4 gpurun_out/sanitize_$tool.log
done
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; tail -14 gpurun_out/pytest_gpu.log
for wl in cfg2 cfg4 cfg3-short; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$wl.json"))
    print("$wl", d["ms_per_step"], d["value"], d.get("e2e", {}).get("value"))
    for k, v in d["kernels"].items(): print("   ", k, v["avg_ms"], v["frac_of_hbm_peak"])
except Exception as e: print("$wl failed", e)
PY
done
