#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/lv_debug.py 8 2>&1 | grep -v "Warning\|^$\|\*\*\*\|OMP_NUM" | tee gpurun_out/lv_debug.txt
nproc; taskset -p $$
