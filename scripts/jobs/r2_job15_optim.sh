#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_optim.py -m gpu -q -x -s --timeout 300 2>&1 | grep -E "worst|Error|error|assert|passed|failed|^E " | tail -25 > gpurun_out/r2_pytest_optim.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_optim.log; cat gpurun_out/r2_pytest_optim.log
