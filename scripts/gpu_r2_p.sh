#!/bin/bash
# round 2, call P (4 GPUs): the driver's scaling command at N=4
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
echo "bench exit $?"; head -c 700 gpurun_out/bench_4gpu.json; echo
