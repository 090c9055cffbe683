#!/bin/bash
# Eight-GPU session: configs[4] (1 Gsps block broadcast one block ahead, 256 of 2048 channels per
# GPU), then config 3 as independent sub-bands (weak scaling) and as one broadcast stream.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi8.txt
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --mode bcast --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu_cfg5.json 2> gpurun_out/bench_8gpu_cfg5.err
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
timeout 600 $TR --master-port 29523 bench.py --gpus 8 --mode bcast --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu_bcast.json 2> gpurun_out/bench_8gpu_bcast.err
for f in bench_8gpu_cfg5 bench_8gpu bench_8gpu_bcast; do head -c 400 gpurun_out/$f.json; echo; tail -c 600 gpurun_out/$f.json; echo; tail -3 gpurun_out/$f.err; done
