#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|flash_attn" -s 186 -c 5 -f -o gpurun_out/prof_mix $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
