#!/bin/bash
# round 2, last evidence pass on one GPU after the bandwidth-kernel session: tests, bench, step profile, ncu launch list + full captures.
# (The reference arm, the DA2 profile, the standalone kernel timings and the pipe microbenchmarks are those of scripts/gpu_final_r2.sh:
#  nothing they measure changed.)  Reports are kept small: gpurun brings back at most 64 MiB.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -n 2 gpurun_out/gpu_tests.log
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; head -c 900 gpurun_out/bench.json; echo; tail -n 3 gpurun_out/bench.err
timeout 300 python scripts/shape_profile_lv.py 2>&1 | grep -v Warning > gpurun_out/shape_profile_lv.txt
python scripts/microbench/rw_mix.py > gpurun_out/rw_mix.txt 2>&1
CMD="python scripts/ncu_step.py"
$CMD > gpurun_out/ncu_plain.log 2>&1 && tail -n 1 gpurun_out/ncu_plain.log
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout -k 10 600 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tc -s 40 -c 6 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm capture exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:flash_attn -s 5 -c 1 -f -o gpurun_out/prof_flash $CMD > gpurun_out/ncu_full_flash.log 2>&1
echo "flash capture exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tail -c 1 -f -o gpurun_out/prof_tail $CMD > gpurun_out/ncu_full_tail.log 2>&1
echo "tail capture exit $?"
timeout -k 10 600 ncu --set full --clock-control none --profile-from-start off -k regex:"layernorm|temporal_attn_tc" -s 20 -c 3 -f -o gpurun_out/prof_mem $CMD > gpurun_out/ncu_full_mem.log 2>&1
echo "mem-bound capture exit $?"
timeout -k 10 600 ncu --set full --clock-control none --profile-from-start off -k regex:"groupnorm|im2col|preprocess|bilinear_slide|window_finalize" -c 12 -f -o gpurun_out/prof_bw $CMD > gpurun_out/ncu_full_bw.log 2>&1
echo "bandwidth-kernel capture exit $?"
du -sh gpurun_out; ls -la gpurun_out/*.ncu-rep
