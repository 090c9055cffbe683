#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 ./tools/bin/membench > gpurun_out/membench.txt 2>&1
python __graft_entry__.py > gpurun_out/build.log 2>&1
B="python bench.py --workload cfg3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/bench_cfg3_t16.json 2> gpurun_out/bench_cfg3_t16.err
RADIOCORE_B200_LIB=$PWD/radio-core_b200/build_t8/libradiocore_b200.so timeout 300 $B > gpurun_out/bench_cfg3_t8.json 2> gpurun_out/bench_cfg3_t8.err
RADIOCORE_B200_LIB=$PWD/radio-core_b200/build_t8/libradiocore_b200.so timeout 300 python tools/gpu_fft_check.py 250000 1000000 10000000 > gpurun_out/fftcheck_t8.txt 2>&1
cat gpurun_out/membench.txt
