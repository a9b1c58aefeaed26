#!/bin/bash
# ncu evidence for bench.py: (1) per-launch durations of one steady-state step, (2) full-section capture of the top GEMM launches.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 161 -c 3 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 30 -c 1 -f -o gpurun_out/prof_flash $CMD > gpurun_out/ncu_full_flash.log 2>&1
echo "flash capture exit $?"
ls -la gpurun_out/
