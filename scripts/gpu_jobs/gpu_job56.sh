#!/bin/bash
# GPU job 56: row-per-warp ROI-align kernel: parity (oracle tests under VBG_ROI_ROW=2) + side-by-side timing at cfg2.
mkdir -p gpurun_out
VBG_ROI_ROW=2 timeout 120 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 100 -k "roi_align or format_aware" 2>&1 | grep -E "passed|failed|FAILED|Error|assert " | tail -8 > gpurun_out/roi_row.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/roi_row.log
timeout 120 python scripts/roi_compare.py cfg2 2>&1 | grep -E "^\[|Error|error" >> gpurun_out/roi_row.log
cat gpurun_out/roi_row.log
