#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "^\[smoke\]" | tail -4
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | grep -E "passed|failed|FAILED|capture failed|^E  " | tail -8 > gpurun_out/r2_pytest_gpu_head.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_gpu_head.log; cat gpurun_out/r2_pytest_gpu_head.log
