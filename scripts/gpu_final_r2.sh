#!/bin/bash
# round 2 end-of-round evidence on one GPU: tests, bench, reference arm, shape profiles, microbenchmarks, ncu launch list + full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -n 2 gpurun_out/gpu_tests.log
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; head -c 900 gpurun_out/bench.json; echo
timeout -k 10 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "reference arm exit $?"; head -c 400 gpurun_out/bench_ref.json; echo
timeout 300 python scripts/shape_profile_lv.py 2>&1 | grep -v Warning > gpurun_out/shape_profile_lv.txt
timeout 300 python scripts/shape_profile_da2.py 2>&1 | grep -v Warning > gpurun_out/shape_profile_da2.txt
(python scripts/run_flash.py; python scripts/run_tail.py) 2>&1 | grep -v Warning > gpurun_out/standalone_kernels.txt
(cd scripts/microbench && for b in pipes tmem mma_rate exp_phase exp_phase2 exp_phase3; do [ -x $b ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../video_depth_normal_v2_b200/csrc -o $b $b.cu -lcuda; echo "== $b"; timeout 60 ./$b; done) > gpurun_out/microbench.txt 2>&1
echo "microbench lines $(wc -l < gpurun_out/microbench.txt)"
CMD="python scripts/ncu_step.py"
$CMD > gpurun_out/ncu_plain.log 2>&1 && tail -n 1 gpurun_out/ncu_plain.log
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc -s 40 -c 6 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm capture exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:flash_attn -s 5 -c 1 -f -o gpurun_out/prof_flash $CMD > gpurun_out/ncu_full_flash.log 2>&1
echo "flash capture exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tail -c 1 -f -o gpurun_out/prof_tail $CMD > gpurun_out/ncu_full_tail.log 2>&1
echo "tail capture exit $?"
timeout -k 10 600 ncu --set full --clock-control none --profile-from-start off -k regex:"layernorm|temporal_attn_tc" -s 20 -c 4 -f -o gpurun_out/prof_mem $CMD > gpurun_out/ncu_full_mem.log 2>&1
echo "mem-bound capture exit $?"
# (the bandwidth-kernel capture lives in scripts/gpu_final_r2b.sh: together with the reports above it exceeds the 64 MiB gpurun brings back)
python scripts/microbench/rw_mix.py > gpurun_out/rw_mix.txt 2>&1; cat gpurun_out/rw_mix.txt
