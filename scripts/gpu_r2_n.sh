#!/bin/bash
# round 2, call N (1 GPU): coalesced register epilogue + 200-register GEMM budget — op tests, RCU conv timings, step profile A/B
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -n 3
python scripts/run_conv.py rcu 2>&1 | grep -v Warning | tee gpurun_out/conv_epilogues_coalesced.txt
VDN_EPI_COALESCED=0 python scripts/run_conv.py "148x148 256->256 +" 2>&1 | grep -v Warning | tee -a gpurun_out/conv_epilogues_coalesced.txt
python scripts/run_gemm.py 2>&1 | grep -v Warning | tail -8
for c in 1 0; do echo "VDN_EPI_COALESCED=$c"; VDN_EPI_COALESCED=$c timeout 300 python scripts/shape_profile_lv.py 2>&1 | grep -v Warning | head -16; done
