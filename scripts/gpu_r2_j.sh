#!/bin/bash
# round 2, call J (1 GPU): ncu full capture of the fused head-tail kernel at the bench shape
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tail -s 8 -c 1 -o gpurun_out/prof_tail -f python scripts/run_tail.py fused > gpurun_out/ncu_tail.log 2>&1
tail -n 5 gpurun_out/ncu_tail.log
