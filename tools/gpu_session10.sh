#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B --workload cfg3 > gpurun_out/bench_cfg3_$name.json 2> gpurun_out/bench_cfg3_$name.err; }
run w0 RC_X=1
run w1 "RC_FFT_SPLIT=256000000:160x160x100x100;1000000:100x100x100;500000:100x100x50"
run w2 "RC_FFT_SPLIT=256000000:128x125x128x125;1000000:200x50x100;500000:125x80x50"
run w3 "RC_FFT_SPLIT=256000000:640x640x625;1000000:250x40x100;500000:250x40x50"
run w4 "RC_FFT_SPLIT=256000000:640x640x625;1000000:125x80x100;500000:200x50x50"
run w5 "RC_FFT_SPLIT=256000000:200x200x80x80;1000000:160x50x125;500000:160x125x25"
timeout 300 $B --workload cfg2 --steps 20 > gpurun_out/bench_cfg2_w.json 2> gpurun_out/bench_cfg2_w.err
timeout 300 $B --workload cfg4 --steps 20 > gpurun_out/bench_cfg4_w.json 2> gpurun_out/bench_cfg4_w.err
