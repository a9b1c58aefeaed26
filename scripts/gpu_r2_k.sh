#!/bin/bash
# round 2, call K (8 GPUs): where do the pinned result pages land and what does it do to the concurrent D2H rate
mkdir -p gpurun_out
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 29517 scripts/d2h_numa_probe.py 2>&1 | grep -v "Warning\|^$\|\*\*\*\|OMP_NUM" | tee gpurun_out/d2h_numa_probe.txt
