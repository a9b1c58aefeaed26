#!/bin/bash
# end-of-round evidence: full gpu test suite, bench (with cpu baseline), reference arm, per-shape profile, ncu launch list + top-kernel captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -n 3 gpurun_out/gpu_tests.log
timeout -k 10 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "reference arm exit $?"; tail -c 600 gpurun_out/bench_ref.json
timeout -k 10 300 python scripts/shape_profile.py > gpurun_out/shape_profile.txt 2>&1
(python scripts/run_gemm.py; python scripts/run_flash.py; python scripts/run_conv.py; python scripts/run_bilinear.py) > gpurun_out/standalone_kernels.txt 2>&1
(cd scripts/microbench && for b in pipes tmem mma_rate exp_phase exp_phase2; do echo "== $b"; timeout 60 ./$b; done) > gpurun_out/microbench.txt 2>&1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --lv-windows 0 --da2-batch 0 --stream-frames 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 290 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 161 -c 4 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm capture exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 30 -c 1 -f -o gpurun_out/prof_flash $CMD > gpurun_out/ncu_full_flash.log 2>&1
echo "flash capture exit $?"
timeout -k 10 600 ncu --set full --clock-control none -k regex:"layernorm|temporal_attn_tc|bilinear" -s 70 -c 6 -f -o gpurun_out/prof_mem $CMD > gpurun_out/ncu_full_mem.log 2>&1
echo "mem-bound capture exit $?"
