#!/bin/bash
# r2 job 10: whole-step CUDA graph of the training path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_graph.py tests/test_gpu_attention_train.py tests/test_gpu_train_step.py -m gpu -q -x --timeout 300 2>&1 | tail -25 > gpurun_out/r2_pytest_train_graph.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_train_graph.log; cat gpurun_out/r2_pytest_train_graph.log
timeout 300 python scripts/train_bench.py cfg2 8 > gpurun_out/r2_train_bench_graph.log 2>&1; echo "train_bench exit $?"
grep -v "^step" gpurun_out/r2_train_bench_graph.log | tail -5
