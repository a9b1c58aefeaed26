#!/bin/bash
# round 2, call M (1 GPU): dwconv7+LN row kernel — op test, DA2 model tests, DA2 shape profile
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -s -k "dwconv" 2>&1 | tail -n 8
timeout -k 10 900 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -k "da2 or DA2 or depth_anything" 2>&1 | tail -n 4
timeout 300 python scripts/shape_profile_da2.py 2>&1 | grep -v Warning > gpurun_out/shape_profile_da2.txt; head -12 gpurun_out/shape_profile_da2.txt
