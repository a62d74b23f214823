#!/bin/bash
# GPU job 24: full ncu capture of the ROI-align kernel only (source-level).
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"roi_align" -c 2 -o gpurun_out/prof_roi3 \
   python scripts/ncu_traffic.py > gpurun_out/ncu_roi3.log 2>&1
tail -3 gpurun_out/ncu_roi3.log
