#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_model_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/model_tests.log 2>&1
echo "model tests exit $?"
grep -E "passed|failed|error|: \{|tap [0-9]|Error|assert" gpurun_out/model_tests.log | tail -40
