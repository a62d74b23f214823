#!/bin/bash
# r2 job 1: ROI-align variants side by side (incl. the persistent forms never run before), ncu --set full of the ROI kernels, baseline bench.
mkdir -p gpurun_out
timeout 120 python scripts/roi_compare.py cfg2 > gpurun_out/r2_roi_compare.log 2>&1; echo "roi_compare exit $?" >> gpurun_out/r2_roi_compare.log
tail -25 gpurun_out/r2_roi_compare.log
VBG_ROI_NCU=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align -c 6 -o gpurun_out/r2_roi_row -f python scripts/roi_compare.py cfg2 > gpurun_out/r2_ncu_roi.log 2>&1; echo "ncu exit $?"
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_base.json 2> gpurun_out/r2_bench_base.err; echo "bench exit $?"
tail -c 3000 gpurun_out/r2_bench_base.json
