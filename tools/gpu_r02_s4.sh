#!/bin/bash
# Round 2, session 4 (1 GPU): full GPU suite (incl. fuzz, unmodified example, short block, graph),
# then the driver's default command lines for both arms.
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; tail -22 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_default.json"))
    print("default", d["ms_per_step"], d["value"], d.get("e2e", {}).get("value"))
    for k in ("wbfm_chain", "short_block", "e2e_dropin", "cpu_baseline"): print(k, d.get(k))
except Exception as e: print("failed", e)
PY
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -3 gpurun_out/bench_reference.err; head -c 1500 gpurun_out/bench_reference.json
