#!/bin/bash
# round 2, call V (1 GPU): straight-line 2x bilinear kernel: op tests, model tests, step profile old / new alternated, ncu
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/v_tests.log 2>&1
echo "gpu tests exit $?"; tail -n 4 gpurun_out/v_tests.log
for i in 1 2; do
  for v in "VDN_BILINEAR_V1=1" "VDN_NONE=1"; do
    echo "== $v"
    env $v timeout 300 python scripts/shape_profile_lv.py 2>&1 | grep -v Warning | grep "total\|groupnorm\|bilinear"
  done
done > gpurun_out/v_profile.txt 2>&1
cat gpurun_out/v_profile.txt
timeout -k 10 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"bilinear" -c 4 -f -o gpurun_out/prof_bil8 python scripts/ncu_step.py > gpurun_out/ncu_full_bil8.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_bil8.ncu-rep
