#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/train_debug.py train_tiny train_tiny_d train_tiny_pre > gpurun_out/job43_debug.log 2>&1
grep -c "^ok" gpurun_out/job43_debug.log; grep -v "^ok" gpurun_out/job43_debug.log | tail -60
