#!/bin/bash
# one full ncu capture of the flash-attention kernel (stand-alone launch at the ViT-L shape) with source-level stall attribution
mkdir -p gpurun_out
export VDN_FA_VARIANT=${VDN_FA_VARIANT:-6}
python scripts/run_flash.py > gpurun_out/fa_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 3 -c 1 -f -o gpurun_out/prof_flash_r2 python scripts/run_flash.py > gpurun_out/fa_ncu.log 2>&1
echo "ncu exit $?"; cat gpurun_out/fa_plain.log; tail -5 gpurun_out/fa_ncu.log
