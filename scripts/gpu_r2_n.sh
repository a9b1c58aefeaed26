#!/bin/bash
# round 2, call N (1 GPU): GEMM kernel variants — op tests, block GEMM and RCU conv timings, step profile
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -n 3
python scripts/run_gemm.py 2>&1 | grep -v Warning | tail -6
python scripts/run_conv.py 2>&1 | grep -v Warning | tee gpurun_out/conv_epilogues_new.txt
timeout 300 python scripts/shape_profile_lv.py 2>&1 | grep -v Warning | head -14
