#!/bin/bash
# GPU job 50: the new kernels / training heads only (CRF NLL forward + backward, training step of the full / crf heads and the
# sampled-loss configuration), with error margins printed.
mkdir -p gpurun_out
VBG_TEST_VERBOSE=1 timeout 200 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_train_step.py -m gpu -q -s --timeout 100 \
  -k "crf_nll or other_heads" 2>&1 | grep -E "MARGIN|passed|failed|FAILED|Error|assert |mismatch|Timeout|gradients off|vbg_" | tail -40 > gpurun_out/pytest_new_heads.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_new_heads.log
cat gpurun_out/pytest_new_heads.log
