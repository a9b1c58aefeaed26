#!/bin/bash
# ncu full-section captures of the top kernels of bench.py (flash attention + the 4 GEMMs of one ViT block + one conv)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 30 -c 1 -f -o gpurun_out/prof_flash $CMD > gpurun_out/ncu_full_flash.log 2>&1
echo "flash capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 161 -c 4 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm capture exit $?"
ls -la gpurun_out/
