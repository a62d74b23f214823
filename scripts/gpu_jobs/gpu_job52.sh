#!/bin/bash
# GPU job 52: attention-backward dK/dV kernel at two CTAs per SM (P / dS alias the transposed Q / dO tiles): parity + training-step profile.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_train_ops.py -m gpu -q --timeout 100 -k "attention_bwd" 2>&1 | grep -E "passed|failed|FAILED|Error|assert " | tail -8 > gpurun_out/pytest_attn_bwd.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_attn_bwd.log; cat gpurun_out/pytest_attn_bwd.log
VBG_TRAIN_PROFILE=1 timeout 200 python scripts/train_bench.py cfg2 6 2>&1 | grep -v -E "Warning|warn|^step|print\(" > gpurun_out/train_step_profile_w.log
head -60 gpurun_out/train_step_profile_w.log | cut -c1-150
