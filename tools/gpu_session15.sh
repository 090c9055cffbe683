#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 120 python tools/gpu_fft_check.py 500000 1000000 2560000 > gpurun_out/fftcheck_fused.txt 2>&1; echo "rc=$?" >> gpurun_out/fftcheck_fused.txt
cat gpurun_out/fftcheck_fused.txt
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { wl=$1; name=$2; shift; shift; env "$@" timeout 200 $B --workload $wl --steps 10 > gpurun_out/bench_${wl}_$name.json 2> gpurun_out/bench_${wl}_$name.err; }
run cfg3 g0 RC_X=1
run cfg3 g1 RC_FUSE_W=64 RC_FUSE_LAG=4 RC_FUSE_NSLOT=8
run cfg3 g2 RC_FUSE_W=128 RC_FUSE_LAG=2 RC_FUSE_NSLOT=5
run cfg3 g3 RC_FUSE_W=256 RC_FUSE_LAG=2 RC_FUSE_NSLOT=4
run cfg3 g4 RC_FUSE_W=256 RC_FUSE_LAG=1 RC_FUSE_NSLOT=3 "RC_FFT_SPLIT=256000000:256x100x100x100"
run cfg3 g5 RC_FUSE_W=32 RC_FUSE_LAG=12 RC_FUSE_NSLOT=24
