#!/bin/bash
# Round 2, session 12 (1 GPU): pass-split sweep for the small transforms; ncu launch lists with DRAM bytes
# (cfg3, cfg3-wbfm); full captures of the angle-storing last IFFT pass and of the first load pass.
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python tools/gpu_split_sweep.py small > gpurun_out/split_sweep_small.txt 2>&1; tail -5 gpurun_out/split_sweep_small.txt
for wl in cfg3 cfg3-wbfm; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  n=$(python -c "import json; print(len(json.load(open('gpurun_out/bench_$wl.json'))['kernels']))")
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
     -k regex:'v3_|fft_pass|ew_kernel|epi_|filtfilt' -s $((3*n)) -c $n --csv --log-file gpurun_out/launches_$wl.csv \
     python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_list_$wl.log 2>&1
done
B1="python bench.py --workload cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'v3_later' -s 10 -c 2 -o /tmp/prof_later $B1 > gpurun_out/ncu_later.log 2>&1
ncu -i /tmp/prof_later.ncu-rep --page raw --csv > gpurun_out/prof_later_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'v3_first' -s 9 -c 1 -o /tmp/prof_first $B1 > gpurun_out/ncu_first.log 2>&1
ncu -i /tmp/prof_first.ncu-rep --page raw --csv > gpurun_out/prof_first_raw.csv 2>/dev/null
du -sh gpurun_out
