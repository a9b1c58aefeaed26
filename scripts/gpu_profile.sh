#!/bin/bash
# ncu evidence for bench.py: (1) per-launch durations of one steady-state step, (2) full-section captures of the top kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --lv-windows 0 --da2-batch 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 161 -c 4 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 30 -c 1 -f -o gpurun_out/prof_flash $CMD > gpurun_out/ncu_full_flash.log 2>&1
echo "flash capture exit $?"
ncu --set full --clock-control none -k regex:"layernorm|temporal_attn_tc|bilinear" -s 70 -c 6 -f -o gpurun_out/prof_mem $CMD > gpurun_out/ncu_full_mem.log 2>&1
echo "mem-bound capture exit $?"
ls -la gpurun_out/*.ncu-rep
