#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/timeline_roi.py > gpurun_out/timeline_roi.log 2>&1; cat gpurun_out/timeline_roi.log
