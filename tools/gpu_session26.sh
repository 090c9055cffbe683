#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { wl=$1; name=$2; shift; shift; env "$@" timeout 300 $B --workload $wl --steps 20 > gpurun_out/bench_${wl}_$name.json 2> gpurun_out/bench_${wl}_$name.err; }
run cfg2 u0 RC_X=1
run cfg2 u1 "RC_FFT_SPLIT=250000:500x500"
run cfg2 u2 "RC_FFT_SPLIT=250000:100x50x50"
run cfg2 u3 "RC_FFT_SPLIT=250000:250x1000"
run cfg4 u0 RC_X=1
run cfg4 u1 "RC_FFT_SPLIT=250000:500x500"
run cfg4 u2 "RC_FFT_SPLIT=250000:100x50x50"
run cfg3 u1 "RC_FFT_SPLIT=1000000:400x50x50"
run cfg3 u2 "RC_FFT_SPLIT=1000000:500x40x50"
run cfg3 u3 "RC_FFT_SPLIT=1000000:250x40x100"
run cfg3 u4 "RC_FFT_SPLIT=1000000:100x100x100"
