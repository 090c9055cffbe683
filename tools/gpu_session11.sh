#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { wl=$1; name=$2; shift; shift; env "$@" timeout 300 $B --workload $wl --steps 10 > gpurun_out/bench_${wl}_$name.json 2> gpurun_out/bench_${wl}_$name.err; }
run cfg3 x0 RC_X=1
run cfg3 x1 "RC_FFT_SPLIT=256000000:160x160x100x100;1000000:160x125x50;500000:160x125x25"
run cfg3 x2 "RC_FFT_SPLIT=256000000:200x160x80x100;1000000:250x80x50;500000:250x40x50"
run cfg3 x3 "RC_FFT_SPLIT=256000000:256x100x100x100;1000000:128x125x... "
run cfg2 x0 RC_X=1
run cfg2 x1 "RC_FFT_SPLIT=10000000:160x250x250;250000:125x50x40;125000:125x40x25;24000:160x150"
run cfg2 x2 "RC_FFT_SPLIT=10000000:200x100x500;250000:125x40x50;125000:250x500;24000:150x160"
run cfg2 x3 "RC_FFT_SPLIT=10000000:250x50x80x10;250000:250x100x10;125000:125x100x10"
run cfg2 x4 "RC_FFT_SPLIT=10000000:250x400x100;250000:200x50x25;125000:250x50x10"
run cfg4 x0 RC_X=1
run cfg4 x1 "RC_FFT_SPLIT=16000000:160x1000x100;250000:125x50x40;125000:125x40x25"
run cfg4 x2 "RC_FFT_SPLIT=16000000:256x250x250;250000:125x40x50;125000:50x50x50"
