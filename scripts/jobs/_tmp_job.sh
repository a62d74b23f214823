#!/bin/bash
mkdir -p gpurun_out
bash scripts/jobs/bench_n.sh 2 2>&1 | grep -v "^\[W\|^W1\|Warning" | cut -c1-1800
timeout 900 python -m pytest tests/test_gpu_reference_scripts.py -m gpu -q -x -s --timeout 900 -k "train" 2>&1 | grep -E "^\[|passed|failed|Error" | tail -8
