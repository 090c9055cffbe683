#!/bin/bash
# One GPU session: parity tests, bench lines, ncu launch list and one full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'v2_|fft_pass|ew_kernel|epilogue' -s 135 -c 90 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'v2_later' -s 20 -c 2 -o gpurun_out/prof_later python bench.py --workload cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_later.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'v2_first' -s 15 -c 3 -o gpurun_out/prof_first python bench.py --workload cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_first.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_cfg3.json | head -c 3000
