#!/bin/bash
# round 2, call I (1 GPU): fused head tail — op tests, A/B bench of the tail, optional model tests
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -s -k "conv_tail or output_conv2" 2>&1 | tail -n 25 | tee gpurun_out/tail_tests.log
timeout 300 python scripts/run_tail.py 2>&1 | grep -v Warning | tee gpurun_out/tail_bench.txt
if [ -n "$MODEL_TESTS" ]; then timeout -k 10 900 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -n 5 | tee -a gpurun_out/tail_tests.log; fi
