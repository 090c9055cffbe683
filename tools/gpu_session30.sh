#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { wl=$1; name=$2; shift; shift; env "$@" timeout 300 $B --workload $wl --steps 10 > gpurun_out/bench_${wl}_$name.json 2> gpurun_out/bench_${wl}_$name.err; }
run cfg4 w0 RC_X=1
run cfg3-wbfm w0 RC_X=1
run cfg2 w0 RC_X=1
