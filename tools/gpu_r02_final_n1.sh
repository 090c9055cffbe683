#!/bin/bash
# Final single-GPU session: smoke(), the driver's two command lines at N = 1.
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time; cat gpurun_out/bench_default.time | grep real
kill $SMI
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("default", d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["roofline_path"]["frac"])
for k in ("wbfm_chain", "short_block", "e2e_dropin", "cpu_baseline"): print(k, str(d.get(k))[:600])
PY
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2> gpurun_out/bench_reference.time; grep real gpurun_out/bench_reference.time; head -c 700 gpurun_out/bench_reference.json
