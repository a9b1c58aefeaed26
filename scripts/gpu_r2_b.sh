#!/bin/bash
# round 2, call B (1 GPU): packed-instruction softmax microbenchmark, per-shape profile of the long-video step, L2-resident head tail
mkdir -p gpurun_out
(cd scripts/microbench && timeout 120 ./exp_phase3) > gpurun_out/exp_phase3.txt 2>&1; cat gpurun_out/exp_phase3.txt
timeout 300 python scripts/shape_profile_lv.py > gpurun_out/shape_profile_lv.txt 2>&1; head -50 gpurun_out/shape_profile_lv.txt
timeout 300 python scripts/l2_tail_experiment.py > gpurun_out/l2_tail.txt 2>&1; cat gpurun_out/l2_tail.txt
