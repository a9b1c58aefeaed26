#!/bin/bash
# round 2, call Q (1 GPU): A/B of two builds of the library in alternation on one box (VDN_LIB_PATH): RCU conv epilogues, block GEMMs, step profile
mkdir -p gpurun_out
BASE=$PWD/video_depth_normal_v2_b200/libvdn_b200_base.so
for i in 1 2; do
  echo "== base"; VDN_LIB_PATH=$BASE python scripts/run_conv.py "148x148 256->256 +" 2>&1 | grep -v Warning
  echo "== new";  python scripts/run_conv.py "148x148 256->256 +" 2>&1 | grep -v Warning
done
echo "== base"; VDN_LIB_PATH=$BASE python scripts/run_gemm.py 2>&1 | grep -v Warning | tail -5
echo "== new";  python scripts/run_gemm.py 2>&1 | grep -v Warning | tail -5
for i in 1 2; do
  echo "== base"; VDN_LIB_PATH=$BASE timeout 300 python scripts/shape_profile_lv.py 2>&1 | grep -v Warning | grep "total\|conv3x3\|flash"
  echo "== new";  timeout 300 python scripts/shape_profile_lv.py 2>&1 | grep -v Warning | grep "total\|conv3x3\|flash"
done
