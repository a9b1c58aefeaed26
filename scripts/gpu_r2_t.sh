#!/bin/bash
# round 2, call T (1 GPU): ncu --set full of the bandwidth kernels of one step (GroupNorm cluster kernel, im2col forms, pre-processing, bilinear)
mkdir -p gpurun_out
CMD="python scripts/ncu_step.py"
timeout -k 10 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"groupnorm|im2col|preprocess|bilinear_slide" -c 14 -f -o gpurun_out/prof_bw $CMD > gpurun_out/ncu_full_bw.log 2>&1
echo "bw capture exit $?"; tail -n 2 gpurun_out/ncu_full_bw.log
python scripts/ncu_summary.py gpurun_out/prof_bw.ncu-rep > gpurun_out/ncu_bw_summary.txt 2>&1; cat gpurun_out/ncu_bw_summary.txt
