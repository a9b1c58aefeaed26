#!/bin/bash
# Debug library with the flash-attention timeline (-DVDN_FA_TIMELINE) next to the product library: VDN_LIB_PATH selects it.
set -e
cd "$(dirname "$0")/../video_depth_normal_v2_b200"
python -m video_depth_normal_v2_b200.build 2>/dev/null || (cd .. && python -m video_depth_normal_v2_b200.build)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DVDN_FA_TIMELINE -c csrc/vdn_attn.cu -o csrc/vdn_attn_tl.o
/usr/local/cuda/bin/nvcc -shared -o libvdn_b200_tl.so csrc/vdn_host.o csrc/vdn_gemm.o csrc/vdn_attn_tl.o csrc/vdn_elem.o csrc/vdn_v5.o csrc/vdn_da2.o csrc/vdn_tail.o -lcudart
echo built libvdn_b200_tl.so
