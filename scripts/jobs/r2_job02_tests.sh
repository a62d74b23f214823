#!/bin/bash
# r2 job 2: new parity tests (headline shapes vs the CPU oracle, live-reference fixtures at cfg2/4/5, mask check, CRF inference, LinearSmall widths)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_headline_parity.py tests/test_gpu_forward.py tests/test_gpu_autograd.py -m gpu -q -x -s --timeout 300 2>&1 | grep -E "^\[|passed|failed|FAILED|Error|error|assert|mismatch|Timeout" | tail -60 > gpurun_out/r2_pytest_parity.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_parity.log; cat gpurun_out/r2_pytest_parity.log
