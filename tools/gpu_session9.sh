#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 > gpurun_out/bench_cfg3_s9.json 2> gpurun_out/bench_cfg3_s9.err
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/bench_cfg2_s9.json 2> gpurun_out/bench_cfg2_s9.err
timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_s9.json 2> gpurun_out/bench_cfg4_s9.err
tail -3 gpurun_out/*_s9.err
