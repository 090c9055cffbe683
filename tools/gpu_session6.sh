#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 300 python tools/gpu_fft_check.py 10000 16000 24000 30000 62500 80000 125000 250000 400000 500000 1000000 2500000 10000000 16000000 > gpurun_out/fftcheck_v3.txt 2>&1
tail -30 gpurun_out/fftcheck_v3.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $B --workload cfg3 > gpurun_out/bench_cfg3_v3.json 2> gpurun_out/bench_cfg3_v3.err
RC_NO_TMA=1 timeout 300 $B --workload cfg3 > gpurun_out/bench_cfg3_v3_notma.json 2> gpurun_out/bench_cfg3_v3_notma.err
RC_FFT_MAXR=800 timeout 300 $B --workload cfg3 > gpurun_out/bench_cfg3_v3_maxr800.json 2> gpurun_out/bench_cfg3_v3_maxr800.err
timeout 300 $B --workload cfg2 --steps 20 > gpurun_out/bench_cfg2_v3.json 2> gpurun_out/bench_cfg2_v3.err
timeout 300 $B --workload cfg4 --steps 20 > gpurun_out/bench_cfg4_v3.json 2> gpurun_out/bench_cfg4_v3.err
B1="python bench.py --workload cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'v3_later' -s 15 -c 2 -o /tmp/prof_later $B1 > gpurun_out/ncu_later.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'v3_first' -s 12 -c 3 -o /tmp/prof_first $B1 > gpurun_out/ncu_first.log 2>&1
for n in later first; do
  ncu -i /tmp/prof_$n.ncu-rep --page raw --csv > gpurun_out/prof_${n}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$n.ncu-rep --page source --csv > gpurun_out/prof_${n}_source.csv 2>/dev/null
done
du -sh gpurun_out
