#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { wl=$1; name=$2; shift; shift; env "$@" timeout 300 $B --workload $wl --steps 10 > gpurun_out/bench_${wl}_$name.json 2> gpurun_out/bench_${wl}_$name.err; }
run cfg3 k0 RADIOCORE_B200_LIB=$PWD/radio-core_b200/build_r200/libradiocore_b200.so
run cfg3 k1 "RC_FFT_SPLIT=256000000:640x800x500"
run cfg3 k2 "RC_FFT_SPLIT=256000000:500x800x640"
run cfg3 k3 "RC_FFT_SPLIT=256000000:640x500x800"
run cfg3 k4 "RC_FFT_SPLIT=256000000:512x500x1000"
run cfg3-wbfm k0 RC_X=1
run cfg2 k0 RADIOCORE_B200_LIB=$PWD/radio-core_b200/build_r200/libradiocore_b200.so
