#!/bin/bash
# round 2, call S (1 GPU): single-launch GroupNorm variants (pass 2 back to front / front to back / two-kernel form), alternated
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x -k "groupnorm or outputs_stay or patch or im2col" > gpurun_out/s_tests.log 2>&1
echo "gpu tests exit $?"; tail -n 2 gpurun_out/s_tests.log
for i in 1 2; do
  for v in "VDN_GN_V1=1" "VDN_GN_FWD=1" "VDN_NONE=1"; do
    echo "== $v"
    env $v timeout 300 python scripts/shape_profile_lv.py 2>&1 | grep -v Warning | grep "total\|groupnorm\|patch_im2col\|bilinear"
  done
done > gpurun_out/s_profile.txt 2>&1
cat gpurun_out/s_profile.txt
