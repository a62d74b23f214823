#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/timeline.py > gpurun_out/timeline2.log 2>&1; grep -A12 "3072x768" gpurun_out/timeline2.log | grep -E "^\[|epilogue|epi_"
