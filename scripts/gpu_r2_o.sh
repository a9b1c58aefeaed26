#!/bin/bash
# round 2, call O (1 GPU): ncu captures of the RCU convolution with the full residual epilogue, per-row and coalesced forms
mkdir -p gpurun_out
VDN_EPI_COALESCED=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 5 -c 1 -f -o gpurun_out/prof_conv_res_v0 python scripts/run_conv.py "+res+res2" > gpurun_out/ncu_conv0.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 5 -c 1 -f -o gpurun_out/prof_conv_res_v1 python scripts/run_conv.py "+res+res2" > gpurun_out/ncu_conv1.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 5 -c 1 -f -o gpurun_out/prof_conv_plain python scripts/run_conv.py "256 plain" > gpurun_out/ncu_conv2.log 2>&1
tail -n 2 gpurun_out/ncu_conv0.log gpurun_out/ncu_conv1.log gpurun_out/ncu_conv2.log
