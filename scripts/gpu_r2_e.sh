#!/bin/bash
# round 2, call E (1 GPU): full gpu test suite, bench
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -n 3 gpurun_out/gpu_tests.log; grep -E "^FAILED|^ERROR" gpurun_out/gpu_tests.log | head -20
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 1500 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
