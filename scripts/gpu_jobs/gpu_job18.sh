#!/bin/bash
# GPU job 18 (2 GPUs): the driver's multi-GPU launch line for both arms.
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r1_n2.json 2> gpurun_out/bench_r1_n2.err; echo "exit $?" >> gpurun_out/bench_r1_n2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r1_n2_ref.json 2> gpurun_out/bench_r1_n2_ref.err; echo "exit $?" >> gpurun_out/bench_r1_n2_ref.err
head -c 1500 gpurun_out/bench_r1_n2.json; echo; tail -3 gpurun_out/bench_r1_n2.err; head -c 400 gpurun_out/bench_r1_n2_ref.json; echo; tail -2 gpurun_out/bench_r1_n2_ref.err
