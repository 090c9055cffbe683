#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
B1="python bench.py --workload cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fused' -s 6 -c 2 -o /tmp/prof_fused $B1 > gpurun_out/ncu_fused.log 2>&1
ncu -i /tmp/prof_fused.ncu-rep --page raw --csv > gpurun_out/prof_fused_raw.csv 2>/dev/null
ncu -i /tmp/prof_fused.ncu-rep --page source --csv > gpurun_out/prof_fused_source.csv 2>/dev/null
tail -3 gpurun_out/ncu_fused.log
