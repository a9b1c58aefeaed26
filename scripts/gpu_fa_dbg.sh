#!/bin/bash
for v in 0 6 1 7 8 10; do echo "== VDN_FA_VARIANT=$v"; VDN_FA_VARIANT=$v python scripts/run_flash.py 2>&1 | tail -1; done | tee gpurun_out/fa_dbg.txt
VDN_FA_VARIANT=6 timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -k "flash_attention" 2>&1 | tail -1
VDN_FA_VARIANT=1 timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -k "flash_attention" 2>&1 | tail -1
