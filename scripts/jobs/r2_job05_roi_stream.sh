#!/bin/bash
# r2 job 5: the TMA row-streaming ROI-align kernel: parity (all variants vs the oracle), side-by-side timing, ncu --set full
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_roi_variants.py tests/test_gpu_ops.py -m gpu -q -x --timeout 120 -k "roi" 2>&1 | tail -15 > gpurun_out/r2_pytest_roi.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_roi.log; cat gpurun_out/r2_pytest_roi.log
for c in cfg2 cfg4 cfg5; do timeout 120 python scripts/roi_compare.py $c >> gpurun_out/r2_roi_stream_compare.log 2>&1; echo "roi_compare $c exit $?" >> gpurun_out/r2_roi_stream_compare.log; done
cat gpurun_out/r2_roi_stream_compare.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align -c 3 -o gpurun_out/r2_roi_stream -f python scripts/roi_compare.py cfg2 ncu > gpurun_out/r2_ncu_roi_stream.log 2>&1; echo "ncu exit $?"
