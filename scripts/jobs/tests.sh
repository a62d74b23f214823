#!/bin/bash
# the -m gpu suite, or the files / -k expression given as arguments
mkdir -p gpurun_out
timeout 1500 python -m pytest ${@:-tests} -m gpu -q -x --timeout 900 2>&1 | grep -E "passed|failed|FAILED|Error|error|assert |^E |mismatch|Timeout" | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
