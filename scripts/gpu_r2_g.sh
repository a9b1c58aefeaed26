#!/bin/bash
# round 2, call G (1 GPU): flash-attention variants at the ViT-L shape + parity of each; hand-off timeline of selected ones
mkdir -p gpurun_out
for v in ${FA_VARIANTS:-6 14 15 16 17 18 19}; do
  echo "== VDN_FA_VARIANT=$v"
  VDN_FA_VARIANT=$v python scripts/run_flash.py 2>&1 | tail -1
  VDN_FA_VARIANT=$v python scripts/run_flash.py 2>&1 | tail -1
  VDN_FA_VARIANT=$v timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -k "flash_attention" 2>&1 | tail -1
done 2>&1 | tee gpurun_out/fa_variants_r2c.txt
for v in ${FA_TL_VARIANTS:-14 15}; do
  echo "== VDN_FA_VARIANT=$v"
  VDN_LIB_PATH=$PWD/video_depth_normal_v2_b200/libvdn_b200_tl.so VDN_FA_VARIANT=$v timeout 120 python scripts/fa_timeline.py 2>&1 | grep -v Warning
done > gpurun_out/fa_timeline_r2c.txt 2>&1
