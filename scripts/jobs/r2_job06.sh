#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/roi_plane_stride.py > gpurun_out/r2_roi_plane_stride.log 2>&1; echo "exit $?" >> gpurun_out/r2_roi_plane_stride.log
cat gpurun_out/r2_roi_plane_stride.log
