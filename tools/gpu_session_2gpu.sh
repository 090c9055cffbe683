#!/bin/bash
# Two-GPU session: one stream broadcast one block ahead (--mode bcast) at config 3, then a short config-5-sized block (N = 1e9) to check memory
# and plan before the 8-GPU run.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --mode bcast --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_bcast.json 2> gpurun_out/bench_2gpu_bcast.err
timeout 900 $TR --master-port 29513 bench.py --gpus 2 --mode bcast --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_2gpu_cfg5.json 2> gpurun_out/bench_2gpu_cfg5.err
for f in bench_2gpu_bcast bench_2gpu_cfg5; do tail -c 700 gpurun_out/$f.json; echo; tail -3 gpurun_out/$f.err; done
