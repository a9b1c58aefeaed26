#!/bin/bash
# round 2, call F (8 GPUs): the driver's scaling command at N=8
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
echo "bench exit $?"; head -c 1200 gpurun_out/bench_8gpu.json; echo; tail -c 1600 gpurun_out/bench_8gpu.json; grep -v "Warning\|^$\|\*\*\*\|OMP_NUM" gpurun_out/bench_8gpu.err | tail -n 8
free -g | head -2; nproc
