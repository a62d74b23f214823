#!/bin/bash
# GPU job 51: roberta-base forward fixture, SyncBatchNorm stage (split backward kernels, one-rank NCCL group), converted-module step.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_forward.py tests/test_gpu_zz_syncbn.py -m gpu -q --timeout 120 -k "rob or syncbn" 2>&1 \
  | grep -E "passed|failed|FAILED|Error|assert |mismatch|Timeout|vbg_|NCCL" | tail -30 > gpurun_out/pytest_rob_syncbn.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_rob_syncbn.log
cat gpurun_out/pytest_rob_syncbn.log
