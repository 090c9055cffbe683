#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
B="python bench.py --workload cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'v2_|fft_pass|ew_kernel|epilogue' -s 30 -c 10 --csv --log-file gpurun_out/launches_cfg3.csv $B > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'v2_later' -s 15 -c 3 -o /tmp/prof_later $B > gpurun_out/ncu_later.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'v2_first' -s 12 -c 3 -o /tmp/prof_first $B > gpurun_out/ncu_first.log 2>&1
for n in later first; do
  ncu -i /tmp/prof_$n.ncu-rep --page raw --csv > gpurun_out/prof_${n}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$n.ncu-rep --page source --csv > gpurun_out/prof_${n}_source.csv 2>/dev/null
  ncu -i /tmp/prof_$n.ncu-rep --page details > gpurun_out/prof_${n}_details.txt 2>/dev/null
  ls -la /tmp/prof_$n.ncu-rep
done
du -sh gpurun_out
