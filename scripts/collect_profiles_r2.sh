#!/bin/bash
# gpurun_out/ of scripts/gpu_final_r2.sh -> the tracked round-2 evidence under profiles/ (run here, after the GPU call came back)
set -e
set +e
O=gpurun_out; P=profiles
cp $O/gpu_tests.log $P/r02_gpu_parity.txt
cp $O/bench.json $P/r02_bench.json
[ -f $O/bench_ref.json ] && [ $O/bench_ref.json -nt $O/bench.json ] && cp $O/bench_ref.json $P/r02_bench_reference_arm.json
cp $O/shape_profile_lv.txt $P/r02_shape_profile_lv.txt
# shape_profile_da2 / standalone_kernels / microbench: only scripts/gpu_final_r2.sh regenerates them (copy by hand after that script)
[ -f $O/rw_mix.txt ] && cp $O/rw_mix.txt $P/r02_rw_mix.txt
cp $O/launches.csv $P/r02_launches.csv
python scripts/summarize_launches.py $O/launches.csv > $P/r02_launches_summary.txt
(for r in prof_gemm prof_flash prof_tail prof_mem prof_bw; do [ -f $O/$r.ncu-rep ] && python scripts/ncu_summary.py $O/$r.ncu-rep && echo; done) > $P/r02_ncu_summary.txt
python scripts/ncu_stalls.py $O/prof_flash.ncu-rep flash_attn > $P/r02_flash_ncu_stalls.txt
python scripts/ncu_stalls.py $O/prof_tail.ncu-rep conv_tail > $P/r02_tail_ncu_stalls.txt
python scripts/ncu_traffic.py $O/prof_gemm.ncu-rep gemm gemm_tc_kernel $P/top_kernel_traffic.json
python scripts/sass_summary.py > $P/r02_sass_summary.txt
echo collected
