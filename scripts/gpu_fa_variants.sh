#!/bin/bash
# flash-attention softmax instruction-mix variants on the real kernel (ViT-L, 32 and 22 frames x 1370 tokens x 16 heads) + parity of each
mkdir -p gpurun_out
for v in 0 1 2 3 4 5 6 7; do
  echo "== VDN_FA_VARIANT=$v"
  VDN_FA_VARIANT=$v python scripts/run_flash.py 2>&1 | tail -3
  VDN_FA_VARIANT=$v timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -k "flash_attention" 2>&1 | tail -1
done 2>&1 | tee gpurun_out/fa_variants.txt
