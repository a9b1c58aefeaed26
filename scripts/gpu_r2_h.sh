#!/bin/bash
# round 2, call H (1 GPU): flash-attention hand-off timeline (debug build with -DVDN_FA_TIMELINE)
mkdir -p gpurun_out
for v in 6 11; do
  echo "== VDN_FA_VARIANT=$v"
  VDN_FA_VARIANT=$v python scripts/fa_timeline.py 2>&1 | grep -v Warning
done > gpurun_out/fa_timeline.txt 2>&1
head -c 6000 gpurun_out/fa_timeline.txt
