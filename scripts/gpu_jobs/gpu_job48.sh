#!/bin/bash
mkdir -p gpurun_out
VBG_PRECISION=fp32 timeout 300 python scripts/train_debug.py train_mid > gpurun_out/job48_debug_mid_fp32.log 2>&1
grep -c "^ok" gpurun_out/job48_debug_mid_fp32.log; grep -v "^ok" gpurun_out/job48_debug_mid_fp32.log | awk '{printf "%s %-70s %s %s %s\n", $1, $2, $3, $4, $5}' | tail -40
