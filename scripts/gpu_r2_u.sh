#!/bin/bash
# round 2, call U (1 GPU): GroupNorm cluster kernel after the instruction diet: full GPU suite, step profile (two-kernel form / new), ncu of the kernel
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/u_tests.log 2>&1
echo "gpu tests exit $?"; tail -n 2 gpurun_out/u_tests.log
for i in 1 2; do
  for v in "VDN_GN_V1=1" "VDN_NONE=1"; do
    echo "== $v"
    env $v timeout 300 python scripts/shape_profile_lv.py 2>&1 | grep -v Warning | grep "total\|groupnorm\|bilinear"
  done
done > gpurun_out/u_profile.txt 2>&1
cat gpurun_out/u_profile.txt
timeout -k 10 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"groupnorm" -c 4 -f -o gpurun_out/prof_gn python scripts/ncu_step.py > gpurun_out/ncu_full_gn.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_gn.ncu-rep
