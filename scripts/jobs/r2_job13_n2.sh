#!/bin/bash
# r2 job 13 (2 GPUs): forward bench and training-step bench at N = 2 (arena all-reduce), reference scripts 2-rank test again with the graphed step
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 10 --warmup 3 --mode train --no-roofline > gpurun_out/r2_bench_train_n2.json 2> gpurun_out/r2_bench_train_n2.err; echo "bench train n2 exit $?"
tail -c 1500 gpurun_out/r2_bench_train_n2.json; tail -3 gpurun_out/r2_bench_train_n2.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --mode train --no-roofline --no-cpu-baseline > gpurun_out/r2_bench_train_n1.json 2> gpurun_out/r2_bench_train_n1.err; echo "bench train n1 exit $?"
tail -c 1200 gpurun_out/r2_bench_train_n1.json
timeout 900 python -m pytest tests/test_gpu_reference_scripts.py -m gpu -q -x -s --timeout 900 -k "train" 2>&1 | grep -E "^\[|passed|failed|Error" | tail -8
