#!/bin/bash
# round 2, call A (1 GPU): box facts, full gpu test suite (incl. the timed-shape parity tests), bench
mkdir -p gpurun_out
(nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv; echo; df -h /dev/shm; free -g; nproc; ulimit -l) > gpurun_out/box_info.txt 2>&1
timeout -k 10 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -n 3 gpurun_out/gpu_tests.log; grep -E "FAILED|Error" gpurun_out/gpu_tests.log | head -20
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 2500 gpurun_out/bench.json; tail -n 15 gpurun_out/bench.err
