#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_train_step.py tests/test_gpu_train_graph.py -m gpu -q -x --timeout 300 2>&1 | grep -E "Error|error|assert|passed|failed|^E " | tail -25 > gpurun_out/r2_pytest_det.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_det.log; cat gpurun_out/r2_pytest_det.log
VBG_TRAIN_PROFILE=1 timeout 300 python scripts/train_bench.py cfg2 5 > gpurun_out/r2_train_profile_b.log 2>&1; echo "train_bench exit $?"
grep -v "^step" gpurun_out/r2_train_profile_b.log | grep -E "cfg2:|roi_align|embed_bwd|profiled" | head
