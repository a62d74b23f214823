#!/bin/bash
mkdir -p gpurun_out
VBG_PRECISION=fp32 timeout 300 python scripts/train_debug.py train_tiny train_tiny_pre > gpurun_out/job44_debug_fp32.log 2>&1
grep -c "^ok" gpurun_out/job44_debug_fp32.log; grep -v "^ok" gpurun_out/job44_debug_fp32.log | awk '{printf "%s %-70s %s %s\n", $1, $2, $3, $4}' | tail -50
