#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
export RC_FFT_SPLIT="256000000:640x640x625;1000000:200x50x100;500000:200x50x50"
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $B --workload cfg3 > gpurun_out/bench_cfg3_pruned.json 2> gpurun_out/bench_cfg3_pruned.err
timeout 600 python -m pytest tests -m gpu -x -q -k "config3 or golden" > gpurun_out/pytest_gpu_part.log 2>&1
B1="python bench.py --workload cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
# v3_first per step: load p0, chan p0 (gather), disc p0, audio p0 -> skip 3 warm-up steps (12), take gather+disc
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'v3_first' -s 13 -c 2 -o /tmp/prof_first $B1 > gpurun_out/ncu_first.log 2>&1
ncu -i /tmp/prof_first.ncu-rep --page raw --csv > gpurun_out/prof_first2_raw.csv 2>/dev/null
ncu -i /tmp/prof_first.ncu-rep --page source --csv > gpurun_out/prof_first2_source.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'filtfilt' -s 3 -c 1 -o /tmp/prof_ff python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ff.log 2>&1
ncu -i /tmp/prof_ff.ncu-rep --page raw --csv > gpurun_out/prof_ff_raw.csv 2>/dev/null
ncu -i /tmp/prof_ff.ncu-rep --page source --csv > gpurun_out/prof_ff_source.csv 2>/dev/null
du -sh gpurun_out
