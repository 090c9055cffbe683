#!/bin/bash
# Last session of the round: GPU parity suite, compute-sanitizer passes over every kernel family,
# then the driver's own command lines (default bench, reference arm).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitize_$tool.log 2>&1
  tail -4 gpurun_out/sanitize_$tool.log
done
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; head -c 500 gpurun_out/bench_default.json
