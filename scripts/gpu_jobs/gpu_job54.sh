#!/bin/bash
# GPU job 54: training step after the host-side fixes (raw stream accessor, asynchronous plan-table upload).
mkdir -p gpurun_out
timeout 200 python scripts/train_bench.py cfg2 8 2>&1 | grep -E "^cfg2|Error|error" > gpurun_out/train_step_x.log
timeout 100 python scripts/train_host_profile.py cfg2 2>&1 | grep -E "host issue" >> gpurun_out/train_step_x.log
cat gpurun_out/train_step_x.log
