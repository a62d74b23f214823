#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_autograd.py -m gpu -q --timeout 100 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
