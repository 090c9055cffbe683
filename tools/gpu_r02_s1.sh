#!/bin/bash
# Round 2, session 1: new parity tests on the GPU, bench lines after the folded pilot filter and
# the (800,640,500) split, ncu captures of the pilot filter before/after.
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
for wl in cfg3 cfg3-wbfm cfg4 cfg2; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$wl.json"))
    print("$wl", d["ms_per_step"], d["value"], d.get("e2e", {}).get("value"))
    for k, v in d["kernels"].items(): print("   ", k, v["avg_ms"], v["frac_of_hbm_peak"])
except Exception as e: print("$wl failed", e)
PY
done
B4="python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
RC_NO_FOLD=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'filtfilt' -s 3 -c 1 -o /tmp/prof_ff_before $B4 > gpurun_out/ncu_ff_before.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'filtfilt' -s 3 -c 1 -o /tmp/prof_ff_after $B4 > gpurun_out/ncu_ff_after.log 2>&1
for w in before after; do
  ncu -i /tmp/prof_ff_$w.ncu-rep --page raw --csv > gpurun_out/prof_ff_${w}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_ff_$w.ncu-rep --page details > gpurun_out/prof_ff_${w}_details.txt 2>/dev/null
done
du -sh gpurun_out
