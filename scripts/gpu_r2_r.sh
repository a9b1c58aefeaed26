#!/bin/bash
# round 2, call R (1 GPU): the bandwidth-kernel pass (single-launch GroupNorm, pixel-form stride-2 im2col, row-form patch im2col,
# identity pre-processing): full GPU suite, step profile old vs new forms (VDN_*_V1 switches), short bench
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r_tests.log 2>&1
echo "gpu tests exit $?"; tail -n 3 gpurun_out/r_tests.log
echo "== old forms" > gpurun_out/r_profile.txt
VDN_GN_V1=1 VDN_IM2COL_V1=1 VDN_PATCH_V1=1 timeout 300 python scripts/shape_profile_lv.py 2>&1 | grep -v Warning | grep "total\|groupnorm\|im2col\|preprocess\|bilinear\|temporal_attn " >> gpurun_out/r_profile.txt
echo "== new forms" >> gpurun_out/r_profile.txt
timeout 300 python scripts/shape_profile_lv.py 2>&1 | grep -v Warning | grep "total\|groupnorm\|im2col\|preprocess\|bilinear\|temporal_attn " >> gpurun_out/r_profile.txt
cat gpurun_out/r_profile.txt
timeout -k 10 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
echo "bench exit $?"; head -c 1200 gpurun_out/r_bench.json; echo
