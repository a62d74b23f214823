#!/bin/bash
# r2 job 4 (2 GPUs): torchrun --nproc-per-node 2 train_SROIE.py through the drop-in vs the one-rank run
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_reference_scripts.py -m gpu -q -x -s --timeout 900 -k "two_ranks" 2>&1 | tail -40 > gpurun_out/r2_pytest_scripts_2gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_scripts_2gpu.log; tail -40 gpurun_out/r2_pytest_scripts_2gpu.log
