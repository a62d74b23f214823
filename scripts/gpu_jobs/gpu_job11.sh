#!/bin/bash
# GPU job 11: where does the pre-split GEMM's tile time go?  full ncu capture with source-level stall samples.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ps_kernel -c 8 -o gpurun_out/prof_ps \
   python scripts/ps_probe.py > gpurun_out/ncu_ps.log 2>&1
tail -5 gpurun_out/ncu_ps.log
