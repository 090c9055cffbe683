#!/bin/bash
# Judged artefacts: bench lines (all three single-GPU configs), ncu launch list of one step,
# DRAM traffic per kernel, one full capture of the dominant kernel.  CSV/markdown only.
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python bench.py --workload cfg3 --steps 10 --warmup 3 > gpurun_out/bench_cfg3_final.json 2> gpurun_out/bench_cfg3_final.err
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/bench_cfg2_final.json 2> gpurun_out/bench_cfg2_final.err
timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 3 > gpurun_out/bench_cfg4_final.json 2> gpurun_out/bench_cfg4_final.err
kill $SMI
timeout 600 python bench.py --impl reference --workload cfg3 --steps 2 --warmup 1 > gpurun_out/bench_cfg3_reference.json 2> gpurun_out/bench_cfg3_reference.err
B1="python bench.py --workload cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
# 12 kernels per step; skip the 3 warm-up steps
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'v3_|fft_pass|ew_kernel|epi_|filtfilt' -s 36 -c 12 --csv --log-file gpurun_out/launches_cfg3.csv $B1 > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'v3_first' -s 13 -c 1 -o /tmp/prof_gather $B1 > gpurun_out/ncu_gather.log 2>&1
ncu -i /tmp/prof_gather.ncu-rep --page raw --csv > gpurun_out/prof_gather_raw.csv 2>/dev/null
ncu -i /tmp/prof_gather.ncu-rep --page details > gpurun_out/prof_gather_details.txt 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'v3_later' -s 24 -c 2 -o /tmp/prof_later $B1 > gpurun_out/ncu_later.log 2>&1
ncu -i /tmp/prof_later.ncu-rep --page raw --csv > gpurun_out/prof_later3_raw.csv 2>/dev/null
ncu -i /tmp/prof_later.ncu-rep --page details > gpurun_out/prof_later3_details.txt 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
du -sh gpurun_out
