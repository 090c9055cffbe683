#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -c 1200 gpurun_out/bench_8gpu.json; tail -5 gpurun_out/bench_8gpu.err
