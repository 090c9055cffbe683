#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 120 python tools/gpu_fft_check.py 500000 1000000 2560000 800000 > gpurun_out/fftcheck_fused.txt 2>&1; echo "rc=$?" >> gpurun_out/fftcheck_fused.txt
cat gpurun_out/fftcheck_fused.txt
RC_NO_FUSE=1 timeout 120 python tools/gpu_fft_check.py 1000000 > gpurun_out/fftcheck_nofuse.txt 2>&1
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { wl=$1; name=$2; shift; shift; env "$@" timeout 200 $B --workload $wl --steps 10 > gpurun_out/bench_${wl}_$name.json 2> gpurun_out/bench_${wl}_$name.err; }
run cfg3 f0 RC_X=1
run cfg3 f1 RC_FUSE_LAG=3 RC_FUSE_NSLOT=6
run cfg3 f2 RC_FUSE_LAG=10 RC_FUSE_NSLOT=20
run cfg3 f3 "RC_FFT_SPLIT=256000000:256x100x100x100"
run cfg3 f4 "RC_FFT_SPLIT=256000000:160x160x100x100"
run cfg3 nf RC_NO_FUSE=1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
