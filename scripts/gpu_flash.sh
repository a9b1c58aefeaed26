#!/bin/bash
mkdir -p gpurun_out
(cd scripts/microbench && ./tmem) > gpurun_out/tmem.log 2>&1; cat gpurun_out/tmem.log
python scripts/run_flash.py > gpurun_out/flash_time.log 2>&1 && cat gpurun_out/flash_time.log &&
ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 5 -c 1 -f -o gpurun_out/prof_flash2 python scripts/run_flash.py > gpurun_out/ncu_flash2.log 2>&1
echo "ncu exit $?"
