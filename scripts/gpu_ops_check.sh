#!/bin/bash
# Run the GPU op tests in independently time-boxed groups (a hung kernel in one group must not lose the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
for grp in "gemm_plain or gemm_layerscale or gemm_gelu or gemm_geglu or bf16" "conv3x3" "pixel_shuffle or temporal_rowmap or patch_tokens" "qkv_split" "temporal_attention" "layernorm or groupnorm or im2col or bilinear or relu_cast or alignment or sobel or errors"; do
  name=$(echo "$grp" | tr ' ' '_' | cut -c1-40)
  echo "=== group: $grp"
  timeout -k 10 420 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "$grp" -s -p no:cacheprovider > "gpurun_out/ops_${name}.log" 2>&1
  echo "exit $? for $grp"
  tail -n 3 "gpurun_out/ops_${name}.log"
done
