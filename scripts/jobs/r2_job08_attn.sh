#!/bin/bash
# r2 job 8: tcgen05 attention backward + in-kernel attention dropout vs float64 autograd
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention_train.py -m gpu -q -s --timeout 200 2>&1 | grep -E "^\[|passed|failed|FAILED|Error|error|assert|Timeout|trap|illegal" | tail -40 > gpurun_out/r2_pytest_attn_train.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_attn_train.log; cat gpurun_out/r2_pytest_attn_train.log
