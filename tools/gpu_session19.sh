#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 200 python tools/gpu_fft_check.py 10000 24000 250000 500000 1000000 10000000 16000000 > gpurun_out/fftcheck_tw.txt 2>&1; tail -14 gpurun_out/fftcheck_tw.txt
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { wl=$1; name=$2; shift; shift; env "$@" timeout 300 $B --workload $wl --steps 10 > gpurun_out/bench_${wl}_$name.json 2> gpurun_out/bench_${wl}_$name.err; }
run cfg3 n0 RC_X=1
run cfg3-wbfm n0 RC_X=1
run cfg2 n0 RC_X=1
run cfg4 n0 RC_X=1
run cfg3 n1 RC_FUSE=1 RC_FUSE_LAG=12 RC_FUSE_NSLOT=24
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
