#!/bin/bash
# round 2, call L (1 GPU): shape profiles of the long-video step and of the DA2 batch-16 call
mkdir -p gpurun_out
timeout 300 python scripts/shape_profile_lv.py 2>&1 | grep -v Warning > gpurun_out/shape_profile_lv.txt; head -30 gpurun_out/shape_profile_lv.txt
timeout 300 python scripts/shape_profile_da2.py 2>&1 | grep -v Warning > gpurun_out/shape_profile_da2.txt; head -40 gpurun_out/shape_profile_da2.txt
